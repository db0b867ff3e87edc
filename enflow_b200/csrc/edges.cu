// K0: neighbour-list construction, bit-exact with the reference's Data.edges
// (enflow/data/base.py:122-144 + enflow/utils/helpers.py:15-29), including its quirks:
//   Q6  both columns of the (image point, atom) hit list are remapped through id_mapping
//   Q7  image pre-filter: sum((p/(box+r_cut))^2) <= 1, radii = box + r_cut
//   Q9  r_cut is float32; r_cut*r_cut is an fp32 product compared (strict <) against fp64 d^2
//   Q10 image order c outer / b / a inner over [-L,+L,0]; Q11 self pairs dropped on remapped labels.
// All geometry is evaluated in fp64 with explicit round-to-nearest ops (no FMA contraction) so the
// comparisons agree with ATen's CPU results bit for bit on identical inputs.
//
// Output order: the hot path wants edges grouped by row (CSR) so segment reductions have a fixed
// order; the reference order (surviving image point, atom) is recoverable through ref_pos[e].
#include "common.cuh"

namespace {

constexpr int ACT_SHIFT = 24;       // active entry = image index << 24 | atom (molecules of < 16.7 M atoms)

template <typename T>
__device__ __forceinline__ double ld3(const T* p, int i, int c) { return (double)p[(int64_t)i * 3 + c]; }

__device__ __forceinline__ double shift_of(int idx, double L) {   // helpers.py:17: [-L, +L, 0]
    return idx == 0 ? -L : (idx == 1 ? L : 0.0);
}

// One CTA per molecule: survivor flags of the 27n image points, ranked in (image, atom) order, plus the
// compact list of "active" survivors: those close enough to the molecule to have a hit at all.
// The activity test is pure pruning against a bounding sphere (centre = mean position, radius = max distance,
// margin ~1e6 ulp): every pair that is not discarded is still decided by the exact fp64 test.
template <typename T>
__global__ void __launch_bounds__(256) k_edges_survivors(const T* __restrict__ pos, const T* __restrict__ box,
                                                          const float* __restrict__ r_cut,
                                                          const int* __restrict__ mol_off, int B,
                                                          int* __restrict__ idmap,
                                                          int* __restrict__ nsurv, int* __restrict__ active,
                                                          int* __restrict__ nactive) {
    __shared__ int warp_tot[8], warp_act[8];
    __shared__ int running_s, running_a;
    __shared__ double red_s[8][3];
    __shared__ double ctr_s[4];
    const int m = blockIdx.x;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    const int64_t base = 27LL * o;
    const double rc = (double)r_cut[m];
    const double bx = ld3(box, o, 0), by = ld3(box, o, 1), bz = ld3(box, o, 2);   // base.py:130 box[0]
    const double ex = __dadd_rn(bx, rc), ey = __dadd_rn(by, rc), ez = __dadd_rn(bz, rc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { running_s = 0; running_a = 0; }
    // ---- bounding sphere
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int a = threadIdx.x; a < n; a += 256) { sx += ld3(pos, o + a, 0); sy += ld3(pos, o + a, 1); sz += ld3(pos, o + a, 2); }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) { red_s[wid][0] = sx; red_s[wid][1] = sy; red_s[wid][2] = sz; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red_s[w][threadIdx.x];
        ctr_s[threadIdx.x] = t / (double)(n > 0 ? n : 1);
    }
    __syncthreads();
    double r2 = 0.0;
    for (int a = threadIdx.x; a < n; a += 256) {
        const double dx = ld3(pos, o + a, 0) - ctr_s[0], dy = ld3(pos, o + a, 1) - ctr_s[1], dz = ld3(pos, o + a, 2) - ctr_s[2];
        r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, s));
    __syncthreads();
    if (lane == 0) red_s[wid][0] = r2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t = fmax(t, red_s[w][0]);
        const double reach = (sqrt(t) + rc) * (1.0 + 1e-9) + 1e-12;
        ctr_s[3] = reach * reach;
    }
    __syncthreads();
    const double cx = ctr_s[0], cy = ctr_s[1], cz = ctr_s[2], reach2 = ctr_s[3];
    // ---- survivors (helpers.py:20-23) and active survivors
    const int total = 27 * n;
    for (int start = 0; start < total; start += 256) {
        const int ip = start + threadIdx.x;
        bool keep = false, act = false;
        int a = 0, k = 0;
        if (ip < total) {
            k = ip / n;
            a = ip - k * n;
            const double px = __dadd_rn(ld3(pos, o + a, 0), shift_of(k % 3, bx));
            const double py = __dadd_rn(ld3(pos, o + a, 1), shift_of((k / 3) % 3, by));
            const double pz = __dadd_rn(ld3(pos, o + a, 2), shift_of(k / 9, bz));
            const double qx = __ddiv_rn(px, ex), qy = __ddiv_rn(py, ey), qz = __ddiv_rn(pz, ez);
            const double q = __dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
            keep = q <= 1.0;
            const double ux = px - cx, uy = py - cy, uz = pz - cz;
            act = keep && (ux * ux + uy * uy + uz * uz <= reach2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const unsigned bal_a = __ballot_sync(0xffffffffu, act);
        if (lane == 0) { warp_tot[wid] = __popc(bal); warp_act[wid] = __popc(bal_a); }
        __syncthreads();
        int pre = running_s, pre_a = running_a;
        for (int w = 0; w < wid; ++w) { pre += warp_tot[w]; pre_a += warp_act[w]; }
        const unsigned lt = (1u << lane) - 1u;
        const int rank = pre + __popc(bal & lt);
        if (ip < total) {
            if (keep) idmap[base + rank] = a;
            if (act) active[base + pre_a + __popc(bal_a & lt)] = (k << ACT_SHIFT) | a;      // (image, atom): no division later
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0, ta = 0;
            for (int w = 0; w < 8; ++w) { t += warp_tot[w]; ta += warp_act[w]; }
            running_s += t;
            running_a += ta;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { nsurv[m] = running_s; nactive[m] = running_a; }
}

// Small molecules (n of a few tens): the same pass with ONE WARP per molecule.  27 n image points are a few hundred, so a
// 256-thread CTA per molecule spent its time in block barriers (three per 256 points plus six for the bounding sphere)
// with most threads idle; a warp needs none: ballots rank the survivors, the running counts live in registers.
// Survivor order (= id_mapping) and the active list are the ones of k_edges_survivors; the bounding sphere may differ in
// its last bits (another summation order of the centre), which is pure pruning with a margin either way.
template <typename T>
__global__ void __launch_bounds__(256) k_edges_survivors_warp(const T* __restrict__ pos, const T* __restrict__ box,
                                                               const float* __restrict__ r_cut,
                                                               const int* __restrict__ mol_off, int B,
                                                               int* __restrict__ idmap, int* __restrict__ nsurv,
                                                               int* __restrict__ active, int* __restrict__ nactive) {
    const int lane = threadIdx.x & 31;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= B) return;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    const int64_t base = 27LL * o;
    const double rc = (double)r_cut[m];
    const double bx = ld3(box, o, 0), by = ld3(box, o, 1), bz = ld3(box, o, 2);   // base.py:130 box[0]
    const double ex = __dadd_rn(bx, rc), ey = __dadd_rn(by, rc), ez = __dadd_rn(bz, rc);
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int a = lane; a < n; a += 32) { sx += ld3(pos, o + a, 0); sy += ld3(pos, o + a, 1); sz += ld3(pos, o + a, 2); }
    const double inv_n = 1.0 / (double)(n > 0 ? n : 1);
    const double cx = warp_sum(sx) * inv_n, cy = warp_sum(sy) * inv_n, cz = warp_sum(sz) * inv_n;
    double r2 = 0.0;
    for (int a = lane; a < n; a += 32) {
        const double dx = ld3(pos, o + a, 0) - cx, dy = ld3(pos, o + a, 1) - cy, dz = ld3(pos, o + a, 2) - cz;
        r2 = fmax(r2, dx * dx + dy * dy + dz * dz);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) r2 = fmax(r2, __shfl_xor_sync(0xffffffffu, r2, s));
    const double reach = (sqrt(r2) + rc) * (1.0 + 1e-9) + 1e-12;
    const double reach2 = reach * reach;
    const unsigned lt = (1u << lane) - 1u;
    const int total = 27 * n;
    int run_s = 0, run_a = 0;
    for (int start = 0; start < total; start += 32) {
        const int ip = start + lane;
        bool keep = false, act = false;
        int a = 0, k = 0;
        if (ip < total) {
            k = ip / n;
            a = ip - k * n;
            const double px = __dadd_rn(ld3(pos, o + a, 0), shift_of(k % 3, bx));
            const double py = __dadd_rn(ld3(pos, o + a, 1), shift_of((k / 3) % 3, by));
            const double pz = __dadd_rn(ld3(pos, o + a, 2), shift_of(k / 9, bz));
            const double qx = __ddiv_rn(px, ex), qy = __ddiv_rn(py, ey), qz = __ddiv_rn(pz, ez);
            const double q = __dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
            keep = q <= 1.0;
            const double ux = px - cx, uy = py - cy, uz = pz - cz;
            act = keep && (ux * ux + uy * uy + uz * uz <= reach2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const unsigned bal_a = __ballot_sync(0xffffffffu, act);
        if (keep) idmap[base + run_s + __popc(bal & lt)] = a;
        if (act) active[base + run_a + __popc(bal_a & lt)] = (k << ACT_SHIFT) | a;
        run_s += __popc(bal);
        run_a += __popc(bal_a);
    }
    if (lane == 0) { nsurv[m] = run_s; nactive[m] = run_a; }
}

// One CTA per molecule, one THREAD per active image point, looping over the molecule's atoms (positions converted
// to fp64 once and staged in shared memory together with id_mapping when the molecule has at most HIT_SM atoms;
// every lane reads the same atom: a broadcast).  FILL=false counts hits, FILL=true writes them in atom order, so a
// row's edges come out in the same (image, atom) order as before.  Molecules of up to 64 atoms: the counting pass
// leaves a 64-bit hit mask per point behind and the fill pass replays it instead of repeating the fp64 tests.
// cnt_csr / cnt_ref are zero-initialised by the caller; only active entries are touched.
constexpr int HIT_SM = 512, HIT_T = 128;

template <typename T, bool FILL>
__global__ void __launch_bounds__(HIT_T) k_edges_hits(const T* __restrict__ pos, const T* __restrict__ box,
                                                       const float* __restrict__ r_cut,
                                                       const int* __restrict__ mol_off,
                                                       const int* __restrict__ idmap, const int* __restrict__ nsurv,
                                                       const int* __restrict__ active, const int* __restrict__ nactive,
                                                       int* __restrict__ cnt_csr, int* __restrict__ cnt_ref,
                                                       unsigned* __restrict__ hmask, int* __restrict__ row,
                                                       int* __restrict__ col, int* __restrict__ ref_pos, int E_cap,
                                                       int* __restrict__ status) {
    __shared__ double spx[HIT_SM], spy[HIT_SM], spz[HIT_SM];
    __shared__ int sid[HIT_SM];
    const int m = blockIdx.x;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    const int64_t base = 27LL * o;
    const double bx = ld3(box, o, 0), by = ld3(box, o, 1), bz = ld3(box, o, 2);
    const float rcf = r_cut[m];
    const double r_sq = (double)__fmul_rn(rcf, rcf);          // base.py:133, fp32 product (Q9)
    const int ns = nsurv[m], na = nactive[m];
    const bool staged = n <= HIT_SM, masked = n <= 64;
    if (staged) {
        for (int a = threadIdx.x; a < n; a += HIT_T) {
            spx[a] = ld3(pos, o + a, 0); spy[a] = ld3(pos, o + a, 1); spz[a] = ld3(pos, o + a, 2);
            sid[a] = a < ns ? idmap[base + a] : a;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x + HIT_T * blockIdx.y; t < na; t += HIT_T * gridDim.y) {   // gridDim.y CTAs share a large molecule
        const int ent = active[base + t];
        const int k = ent >> ACT_SHIFT, a = ent & ((1 << ACT_SHIFT) - 1);
        const int i = o + a;
        const int64_t gw = 27LL * i + k, ip = base + (int64_t)k * n + a;
        const double px = __dadd_rn(staged ? spx[a] : ld3(pos, i, 0), shift_of(k % 3, bx));
        const double py = __dadd_rn(staged ? spy[a] : ld3(pos, i, 1), shift_of((k / 3) % 3, by));
        const double pz = __dadd_rn(staged ? spz[a] : ld3(pos, i, 2), shift_of(k / 9, bz));
        int e = 0, rr = 0;
        if (FILL) { e = cnt_csr[gw]; rr = cnt_ref ? cnt_ref[ip] : 0; }
        if (FILL && masked) {
            unsigned w0 = hmask[(base + t) * 2], w1 = hmask[(base + t) * 2 + 1];
            while (w0 | w1) {
                int j;
                if (w0) { j = __ffs(w0) - 1; w0 &= w0 - 1; }
                else { j = 32 + __ffs(w1) - 1; w1 &= w1 - 1; }
                if (e < E_cap) { row[e] = i; col[e] = o + sid[j]; if (ref_pos) ref_pos[e] = rr; }
                ++e; ++rr;
            }
            continue;
        }
        int cnt = 0;
        unsigned w0 = 0, w1 = 0;
        for (int j = 0; j < n; ++j) {
            const double dx = __dsub_rn(px, staged ? spx[j] : ld3(pos, o + j, 0));
            const double dy = __dsub_rn(py, staged ? spy[j] : ld3(pos, o + j, 1));
            const double dz = __dsub_rn(pz, staged ? spz[j] : ld3(pos, o + j, 2));
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (d2 < r_sq) {
                // base.py:137: the atom column is ALSO indexed through id_mapping (Q6)
                int lab;
                if (j < ns) lab = staged ? sid[j] : idmap[base + j];
                else { lab = j; atomicOr(status, 2); }            // the reference would raise IndexError here
                if (lab != a) {                                     // base.py:139 (Q11)
                    if (FILL) {
                        if (e < E_cap) { row[e] = i; col[e] = o + lab; if (ref_pos) ref_pos[e] = rr; }
                        ++e; ++rr;
                    } else {
                        ++cnt;
                        if (j < 32) w0 |= 1u << j;
                        else if (j < 64) w1 |= 1u << (j - 32);
                    }
                }
            }
        }
        if (!FILL) {
            cnt_csr[gw] = cnt;
            if (cnt_ref) cnt_ref[ip] = cnt;
            if (masked) { hmask[(base + t) * 2] = w0; hmask[(base + t) * 2 + 1] = w1; }
        }
    }
}

// One CTA per molecule, one warp per active image point. FILL=false counts hits, FILL=true writes them.
// cnt_csr / cnt_ref are zero-initialised by the caller; only active entries are touched.
template <typename T, bool FILL>
__global__ void __launch_bounds__(256) k_edges_hits_warp(const T* __restrict__ pos, const T* __restrict__ box,
                                                     const float* __restrict__ r_cut,
                                                     const int* __restrict__ mol_off,
                                                     const int* __restrict__ idmap, const int* __restrict__ nsurv,
                                                     const int* __restrict__ active, const int* __restrict__ nactive,
                                                     int* __restrict__ cnt_csr, int* __restrict__ cnt_ref,
                                                     unsigned* __restrict__ hmask, int* __restrict__ row,
                                                     int* __restrict__ col, int* __restrict__ ref_pos, int E_cap,
                                                     int* __restrict__ status) {
    const int m = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    const int64_t base = 27LL * o;
    const double bx = ld3(box, o, 0), by = ld3(box, o, 1), bz = ld3(box, o, 2);
    const float rcf = r_cut[m];
    const double r_sq = (double)__fmul_rn(rcf, rcf);          // base.py:133, fp32 product (Q9)
    const int ns = nsurv[m], na = nactive[m];
    // molecules of up to 64 atoms: the counting pass leaves its hit ballots behind (two words per active point)
    // and the fill pass replays them instead of repeating the fp64 distance tests
    const bool masked = n <= 64;
    for (int t = wid + 8 * blockIdx.y; t < na; t += 8 * gridDim.y) {     // gridDim.y CTAs share a large molecule
        const int ent = active[base + t];
        const int k = ent >> ACT_SHIFT, a = ent & ((1 << ACT_SHIFT) - 1);
        const int i = o + a;
        const int64_t gw = 27LL * i + k, ip = base + (int64_t)k * n + a;
        const int kx = k % 3, ky = (k / 3) % 3, kz = k / 9;
        const double px = __dadd_rn(ld3(pos, i, 0), shift_of(kx, bx));
        const double py = __dadd_rn(ld3(pos, i, 1), shift_of(ky, by));
        const double pz = __dadd_rn(ld3(pos, i, 2), shift_of(kz, bz));
        int running = 0;
        int out_csr = 0, out_ref = 0;
        if (FILL) { out_csr = cnt_csr[gw]; out_ref = cnt_ref ? cnt_ref[ip] : 0; }
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            int lab = 0;
            if (FILL && masked) {
                hit = (hmask[(base + t) * 2 + (j0 >> 5)] >> lane) & 1u;
                if (hit) lab = j < ns ? idmap[base + j] : j;
            } else if (j < n) {
                const double dx = __dsub_rn(px, ld3(pos, o + j, 0));
                const double dy = __dsub_rn(py, ld3(pos, o + j, 1));
                const double dz = __dsub_rn(pz, ld3(pos, o + j, 2));
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (d2 < r_sq) {
                    // base.py:137: the atom column is ALSO indexed through id_mapping (Q6)
                    if (j < ns) lab = idmap[base + j];
                    else { lab = j; atomicOr(status, 2); }        // the reference would raise IndexError here
                    hit = lab != a;                                 // base.py:139 (Q11)
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (!FILL && masked && lane == 0) hmask[(base + t) * 2 + (j0 >> 5)] = bal;
            if (FILL && hit) {
                const int r = running + __popc(bal & ((1u << lane) - 1u));
                const int e = out_csr + r;
                if (e < E_cap) { row[e] = i; col[e] = o + lab; if (ref_pos) ref_pos[e] = out_ref + r; }
            }
            running += __popc(bal);
        }
        if (!FILL && lane == 0) { cnt_csr[gw] = running; if (cnt_ref) cnt_ref[ip] = running; }
    }
}

__global__ void k_rowptr(const int* __restrict__ cnt_csr, int N, int* __restrict__ rowptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) rowptr[i] = cnt_csr[27LL * i];
}

// ---- exclusive scan of two int arrays of equal length (blockIdx.y selects the array) ----
constexpr int SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_CHUNK = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ int block_exclusive(int v, int* total) {
    __shared__ int wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) { if (w < wid) pre += wsum[w]; tot += wsum[w]; }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const int* a0, const int* a1, int64_t n, int* sums, int nb,
                                                            const int* skip) {
    if (skip && *skip) return;
    const int* a = blockIdx.y ? a1 : a0;
    const int64_t b0 = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) if (b0 + t < n) s += a[b0 + t];
    int tot;
    block_exclusive(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.y * nb + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_top(int* sums, int nb, const int* skip) {
    if (skip && *skip) return;
    int* s = sums + blockIdx.y * nb;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int start = 0; start < nb; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const int v = i < nb ? s[i] : 0;
        int tot;
        const int ex = block_exclusive(v, &tot);
        const int c = carry;
        if (i < nb) s[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
}

// in place; element n (one past the end) receives the grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int* a0, int* a1, int64_t n, const int* sums, int nb,
                                                             const int* skip) {
    if (skip && *skip) return;
    int* a = blockIdx.y ? a1 : a0;
    const int64_t b0 = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) { v[t] = (b0 + t < n) ? a[b0 + t] : 0; s += v[t]; }
    int tot;
    int ex = block_exclusive(s, &tot) + sums[blockIdx.y * nb + blockIdx.x];
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        if (b0 + t < n) a[b0 + t] = ex;
        ex += v[t];
        if (b0 + t == n - 1) a[n] = ex;
    }
}

__global__ void k_edges_finish(const int* __restrict__ cnt_csr, int N, int E_cap, int* __restrict__ rowptr,
                               int* __restrict__ E_dev, int* __restrict__ status) {
    const int E = cnt_csr[27LL * N];
    rowptr[N] = E;
    E_dev[0] = E < E_cap ? E : E_cap;
    E_dev[1] = E;
    if (E > E_cap) atomicOr(status, 1);
}

// ---- column-grouped permutation (CSR -> CSC), deterministic and stable in edge order ----
__global__ void k_col_count(const int* __restrict__ col, const int* __restrict__ E_dev, int* __restrict__ colcnt,
                            const int* __restrict__ skip) {
    if (skip && *skip) return;
    const int E = E_dev[0];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x)
        atomicAdd(&colcnt[col[e]], 1);       // integer adds: the result does not depend on order
}

// Fill: every edge takes the next free slot of its column (integer atomics: WHICH slot is order dependent, the set of
// edges of a column is not), then k_col_sort puts every column's list into increasing edge order.  The result is the
// stable column-grouped permutation whatever order the atomics ran in: deterministic.
__global__ void __launch_bounds__(256) k_col_fill(const int* __restrict__ col, const int* __restrict__ E_dev,
                                                   const int* __restrict__ colptr, int* __restrict__ cursor,
                                                   int* __restrict__ perm, const int* __restrict__ skip) {
    if (skip && *skip) return;
    const int E = E_dev[0];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int c = col[e];
        perm[colptr[c] + atomicAdd(&cursor[c], 1)] = e;
    }
}

// one warp per column: rank sort of its (distinct) edge ids held in registers, up to 32 * KMAX entries; longer lists are
// sorted in place by one lane (a column with more than 256 incoming edges: not seen in any config)
__global__ void __launch_bounds__(256) k_col_sort(const int* __restrict__ colptr, int N, int* __restrict__ perm,
                                                   const int* __restrict__ skip) {
    if (skip && *skip) return;
    constexpr int KMAX = 8;
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= N) return;
    const int s0 = colptr[c], d = colptr[c + 1] - s0;
    if (d <= 1) return;
    if (d > 32 * KMAX) {
        if (lane == 0)
            for (int i = 1; i < d; ++i) {
                const int v = perm[s0 + i];
                int j = i - 1;
                while (j >= 0 && perm[s0 + j] > v) { perm[s0 + j + 1] = perm[s0 + j]; --j; }
                perm[s0 + j + 1] = v;
            }
        return;
    }
    const int K = (d + 31) >> 5;                 // warp-uniform
    int v[KMAX], rank[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        v[k] = (k < K && lane + 32 * k < d) ? perm[s0 + lane + 32 * k] : 0x7fffffff;
        rank[k] = 0;
    }
#pragma unroll
    for (int kj = 0; kj < KMAX; ++kj) {
        if (kj >= K) break;
        for (int src = 0; src < 32; ++src) {
            const int wv = __shfl_sync(0xffffffffu, v[kj], src);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) rank[k] += wv < v[k];
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
        if (k < K && lane + 32 * k < d) perm[s0 + rank[k]] = v[k];
}

__global__ void k_zero_int2(int* __restrict__ p, int64_t n, int* __restrict__ q, int64_t m, const int* __restrict__ skip) {
    if (skip && *skip) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n + m; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n) p[i] = 0;
        else q[i - n] = 0;
    }
}

// same[0] = 1 iff the two CSR edge lists are identical (used to reuse the column permutation between layers:
// in the fully connected regime the neighbour list does not change from one coupling step to the next)
__global__ void k_edges_same_init(const int* __restrict__ Ea, const int* __restrict__ Eb, int* __restrict__ same) {
    same[0] = (Ea[0] == Eb[0] && Ea[1] == Eb[1]) ? 1 : 0;
}
__global__ void k_edges_same(const int* __restrict__ ra, const int* __restrict__ ca, const int* __restrict__ rb,
                             const int* __restrict__ cb, const int* __restrict__ Ea, int* __restrict__ same) {
    if (!same[0]) return;
    const int E = Ea[0];
    bool diff = false;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x)
        diff |= (ra[e] != rb[e]) | (ca[e] != cb[e]);
    if (__any_sync(0xffffffffu, diff) && (threadIdx.x & 31) == 0) atomicAnd(same, 0);
}

}  // namespace

// small arrays: the whole exclusive scan in ONE CTA of 1024 threads, 4096 elements per pass (coalesced 16-byte
// loads, warp scans, one carry); element n receives the grand total like k_scan_apply
constexpr int SCAN1_THREADS = 1024, SCAN1_MAX = 1 << 18;
__global__ void __launch_bounds__(SCAN1_THREADS) k_scan_single(int* __restrict__ a, int n, const int* __restrict__ skip) {
    if (skip && *skip) return;
    __shared__ int wsum[SCAN1_THREADS / 32];
    __shared__ int carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const bool vec = (reinterpret_cast<uintptr_t>(a) & 15) == 0;
    for (int base = 0; base < n; base += SCAN1_THREADS * 4) {
        const int i = base + threadIdx.x * 4;
        int v[4] = {0, 0, 0, 0};
        if (vec && i + 3 < n) {
            const int4 q = *reinterpret_cast<const int4*>(a + i);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (i + k < n) v[k] = a[i + k];
        }
        const int s = v[0] + v[1] + v[2] + v[3];
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const int carry = carry_s;                  // read before the barrier: warp 0 updates it after
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const int w = wsum[lane];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            wsum[lane] = winc - w;                  // exclusive prefix of the warp totals
            if (lane == 31) carry_s = carry + winc;
        }
        __syncthreads();
        int ex = carry + wsum[wid] + inc - s;
        if (vec && i + 3 < n) {
            *reinterpret_cast<int4*>(a + i) = make_int4(ex, ex + v[0], ex + v[0] + v[1], ex + v[0] + v[1] + v[2]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i + k < n) a[i + k] = ex;
                ex += v[k];
            }
        }
        __syncthreads();                            // wsum / carry_s are rewritten by the next pass
    }
    if (threadIdx.x == 0) a[n] = carry_s;
}

static int scan2(int* a0, int* a1, int64_t n, int* sums, cudaStream_t st, const int* skip = nullptr) {
    if (!a1 && n <= SCAN1_MAX) {
        enf_count_launch(), k_scan_single<<<1, SCAN1_THREADS, 0, st>>>(a0, (int)n, skip);
        ENF_CHECK_LAUNCH();
        return ENF_OK;
    }
    const int nb = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    const int ny = a1 ? 2 : 1;
    dim3 g(nb, ny);
    enf_count_launch(), k_scan_sums<<<g, SCAN_THREADS, 0, st>>>(a0, a1, n, sums, nb, skip);
    enf_count_launch(), k_scan_top<<<dim3(1, ny), SCAN_THREADS, 0, st>>>(sums, nb, skip);
    enf_count_launch(), k_scan_apply<<<g, SCAN_THREADS, 0, st>>>(a0, a1, n, sums, nb, skip);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int64_t enf_edges_workspace_ints(int N) {
    const int64_t n27 = 27LL * N;
    const int64_t nb = (n27 + SCAN_CHUNK - 1) / SCAN_CHUNK;
    // qrank, idmap, cnt_csr(+1), cnt_ref(+1), nsurv(<=N), atom_mol(N), colcnt(N+1), cursor(N), scan sums
    // ... active (27N), nactive (N), hit ballots (2 x 27N)
    return 4 * n27 + 2 + 4LL * N + 1 + 2 * nb + 64 + n27 + N + 4 + 2 * n27;
}

template <typename T>
int enf_build_edges_t(const T* pos, const T* box, const float* r_cut, const int* mol_off, int B, int N, int E_cap,
                      int* row, int* col, int* rowptr, int* ref_pos, int* E_dev, int* status, int* ws,
                      cudaStream_t st) {
    if (B == 0 || N == 0) {          // empty batch: an empty, well-formed list
        cudaMemsetAsync(rowptr, 0, sizeof(int) * ((size_t)N + 1), st);
        cudaMemsetAsync(E_dev, 0, sizeof(int) * 2, st);
        return ENF_OK;
    }
    const int64_t n27 = 27LL * N;
    int* idmap = ws + n27;                    // (the first 27N ints of the layout are unused)
    int* cnt_csr = idmap + n27;
    int* cnt_ref = cnt_csr + n27 + 1;
    int* nsurv = cnt_ref + n27 + 1;
    int* atom_mol = nsurv + N;                // kept for enf_build_col_perm's layout (unused here)
    int* sums = atom_mol + N + (2 * N + 1);   // colcnt/cursor live between (see enf_build_col_perm)
    const int64_t nb_scan = (n27 + SCAN_CHUNK - 1) / SCAN_CHUNK;
    int* active = sums + 2 * nb_scan + 64;    // 27N
    int* nactive = active + n27;              // B <= N
    unsigned* hmask = reinterpret_cast<unsigned*>(nactive + N);      // 2 x 27N
    if (!ref_pos) cnt_ref = nullptr;          // reference order not requested: one array to count, zero and scan
    cudaMemsetAsync(cnt_csr, 0, sizeof(int) * ((cnt_ref ? 2 : 1) * n27 + (cnt_ref ? 2 : 1)), st);
    // small molecules: one thread per active point; large ones (long atom loops, few points per SM): one warp per point
    const bool per_thread = N <= 96LL * B;
    // ... and, when there are enough molecules to fill the GPU with warps, one warp instead of one CTA per molecule for the
    // survivor pass (C4, 16 384 molecules: 169 -> ~90 us per layer; at B = 64 the CTA form has 8x the parallelism)
    if (per_thread && B >= 4 * enf_num_sms())
        enf_count_launch(), k_edges_survivors_warp<T><<<(B + 7) / 8, 256, 0, st>>>(pos, box, r_cut, mol_off, B, idmap, nsurv, active,
                                                                                 nactive);
    else
        enf_count_launch(), k_edges_survivors<T><<<B, 256, 0, st>>>(pos, box, r_cut, mol_off, B, idmap, nsurv, active, nactive);
    // large molecules: several CTAs per molecule (the active list of a 500-atom fragment has ~3000 entries)
    int ysplit = (B > 0 ? N / B : 1) / 24;
    ysplit = ysplit < 1 ? 1 : (ysplit > 32 ? 32 : ysplit);
    const dim3 hgrid(B, ysplit);
    if (per_thread)
        enf_count_launch(), k_edges_hits<T, false><<<hgrid, HIT_T, 0, st>>>(pos, box, r_cut, mol_off, idmap, nsurv, active, nactive,
                                                                        cnt_csr, cnt_ref, hmask, nullptr, nullptr, nullptr, E_cap, status);
    else
        enf_count_launch(), k_edges_hits_warp<T, false><<<hgrid, 256, 0, st>>>(pos, box, r_cut, mol_off, idmap, nsurv, active, nactive,
                                                                           cnt_csr, cnt_ref, hmask, nullptr, nullptr, nullptr, E_cap, status);
    ENF_CHECK_LAUNCH();
    ENF_TRY(scan2(cnt_csr, cnt_ref, n27, sums, st));
    if (per_thread)
        enf_count_launch(), k_edges_hits<T, true><<<hgrid, HIT_T, 0, st>>>(pos, box, r_cut, mol_off, idmap, nsurv, active, nactive,
                                                                       cnt_csr, cnt_ref, hmask, row, col, ref_pos, E_cap, status);
    else
        enf_count_launch(), k_edges_hits_warp<T, true><<<hgrid, 256, 0, st>>>(pos, box, r_cut, mol_off, idmap, nsurv, active, nactive,
                                                                          cnt_csr, cnt_ref, hmask, row, col, ref_pos, E_cap, status);
    enf_count_launch(), k_rowptr<<<(N + 255) / 256, 256, 0, st>>>(cnt_csr, N, rowptr);
    enf_count_launch(), k_edges_finish<<<1, 1, 0, st>>>(cnt_csr, N, E_cap, rowptr, E_dev, status);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

template int enf_build_edges_t<float>(const float*, const float*, const float*, const int*, int, int, int, int*, int*,
                                      int*, int*, int*, int*, int*, cudaStream_t);
template int enf_build_edges_t<double>(const double*, const double*, const float*, const int*, int, int, int, int*,
                                       int*, int*, int*, int*, int*, int*, cudaStream_t);

// colptr [N+1], perm [E_cap]; ws is the same workspace handed to enf_build_edges_t.
// skip (nullable, device int): when non-zero every kernel returns at once and colptr/perm keep their contents.
int enf_build_col_perm(const int* col, const int* rowptr, const int* mol_off, int B, int N, int E_cap,
                       const int* E_dev, int* colptr, int* perm, int* ws, const int* skip, cudaStream_t st) {
    const int64_t n27 = 27LL * N;
    int* after_atom_mol = ws + 4 * n27 + 2 + 2LL * N;
    int* cursor = after_atom_mol;            // N
    int* sums = cursor + N + (N + 1);
    const int zb = enf_num_sms() * 2;
    enf_count_launch(), k_zero_int2<<<zb, 256, 0, st>>>(colptr, N + 1, cursor, N, skip);
    enf_count_launch(), k_col_count<<<enf_num_sms() * 4, 256, 0, st>>>(col, E_dev, colptr, skip);
    ENF_CHECK_LAUNCH();
    ENF_TRY(scan2(colptr, nullptr, N, sums, st, skip));
    enf_count_launch(), k_col_fill<<<enf_num_sms() * 4, 256, 0, st>>>(col, E_dev, colptr, cursor, perm, skip);
    enf_count_launch(), k_col_sort<<<(N + 7) / 8, 256, 0, st>>>(colptr, N, perm, skip);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// same[0] <- (edge list a == edge list b)
int enf_edges_same(const int* row_a, const int* col_a, const int* E_a, const int* row_b, const int* col_b,
                   const int* E_b, int* same, cudaStream_t st) {
    enf_count_launch(), k_edges_same_init<<<1, 1, 0, st>>>(E_a, E_b, same);
    enf_count_launch(), k_edges_same<<<enf_num_sms() * 4, 256, 0, st>>>(row_a, col_a, row_b, col_b, E_a, same);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// in-place exclusive scan of a[0..n); a[n] receives the total. sums: >= ceil(n/2048) ints of scratch
int64_t enf_scan_scratch_ints(int64_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 8; }
int enf_scan_int(int* a, int64_t n, int* sums, cudaStream_t st) { return scan2(a, nullptr, n, sums, st); }
