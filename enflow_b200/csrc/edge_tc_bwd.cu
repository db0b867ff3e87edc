// K1 backward on tcgen05: recompute + dgrad + wgrad of the two dense edge layers, per 64-edge tile, without
// reading any saved [E,H] activation (enflow/nn/egcl.py:57-63,71-75 differentiated by hand).
//
// All GEMMs keep the accumulator TRANSPOSED (TMEM lane = hidden unit, column = edge or weight column):
//   G1  T [n][e] = W2 x1^T            A = W2 image (K-major)           B = x1^T image [k][e] (MN-major)
//   G2  T [n][e] = W3 x2^T            A = W3 image (K-major)           B = x2^T image [k][e] (MN-major)
//   G3  T [k][e] = W3^T dz3^T         A = W3 image (MN-major)          B = dz3^T image [n][e] (MN-major)
//       TW3[n][k] += dz3^T x2         A = dz3^T image [n][e] (K-major) B = x2^T image [k][e] (K-major)
//   G4  T [k][e] = W2^T dz2^T         A = W2 image (MN-major)          B = dz2^T image [n][e] (MN-major)
//       TW2[n][k] += dz2^T x1         A = dz2^T image [n][e] (K-major) B = x1^T image [k][e] (K-major)
// One swizzled image per operand serves every view (tc_common.cuh).  TW2/TW3 stay in TMEM for the life of the CTA and
// are written once at the end as a per-CTA partial; partials are combined in CTA order (deterministic).
//
// Structure (round 2): WARP-SPECIALISED, TWO TILES IN FLIGHT.  Per tile the chain is strictly serial
// (x1 -> G1 -> E1 -> G2 -> E2 -> G3 -> E3 -> G4 -> E4; Ek = epilogue that turns an accumulator into the next operand), so
// one tile alone leaves the tensor pipe idle during every epilogue and the CUDA cores idle during every MMA
// (profiles/r1c_phase_times.txt: 14.9 k cycles per tile, 7.3 k of MMA, ~9 k of epilogue).  Now:
//   * warp 16 issues every tcgen05.mma (one elected lane) and the TMA copies; warps 0..15 only run epilogues.
//     They never meet at a block barrier inside the tile loop: an epilogue task ends with mbarrier arrivals, the
//     issuing warp waits for all 512, issues the next GEMM of that tile and commits to the tile's accumulator barrier.
//   * a CTA's tiles alternate between two pipelines (A = even, B = odd), each with its own accumulator (64 TMEM
//     columns), its own x buffer and its own barriers.  Every epilogue warp runs the fixed task order
//        E2(a) E4(b')+X(b) E3(a) E1(b) E4(a)+X(a') E2(b) E1(a') E3(b) | ...
//     (a, b = current tiles of A and B, ' = previous/next tile of that pipeline, X = gather + x1 image), so the GEMM a
//     task waits for was issued one task earlier and ran underneath the other pipeline's epilogue.
//   * shared memory (fp32-accurate mode): 128 KB resident weight images + x_A, x_B, and ONE dz buffer (32 KB each,
//     hi + lo): a tile needs the dz buffer only from E2 to the end of G4, and the task order above keeps those
//     windows of A and B disjoint except for one GEMM (G4(b) in front of E2(a')), which is the one exposed wait per
//     two tiles.  bf16 mode has room for a dz buffer per pipeline (no exposed wait).
//   * silu'(z2) of a tile (needed two tasks later) is parked in 64 spare TMEM columns per pipeline instead of
//     registers (tcgen05.st / tcgen05.ld), which is what lets two tiles' state fit in the ~100 registers per thread that 20 warps leave.
//
// Thread map of the epilogue warps: warp w owns TMEM lanes [32 (w%4), +32) = hidden units n and tile edges
// [16 (w/4), +16).  Per-tile edge data (k_edge_geom_bwd): row/col indices arrive in a 3-slot shared-memory ring by
// TMA one tile ahead; the other per-edge scalars (r, the force-branch seed ds, wrapped differences, run headers) are
// read from global memory with warp-uniform vector loads where they are used.
#include "common.cuh"
#include "tc_common.cuh"

// Debug builds (make PHASE=1): clock64 stamps of one epilogue thread (3 per task: start, accumulator ready, done) and of
// the MMA-issuing lane (2 per GEMM: operands ready, issued) over four periods of one CTA, read back by
// tools/phase_times.py.  Not part of the product library (the extra symbol is not in include/enflow_b200.h).
#ifdef ENF_PHASE_TIMING
__device__ long long g_phase[2 * 4 * 32];
#define PH_ON(k) (blockIdx.x == 3 && (k) >= 20 && (k) < 24)
#define STAMP_E(k, i) do { if (PH_ON(k) && tid == 160) g_phase[((k) - 20) * 32 + (i)] = clock64(); } while (0)
#define STAMP_M(k, i) do { if (PH_ON(k) && lane == 0) g_phase[128 + ((k) - 20) * 32 + (i)] = clock64(); } while (0)
#pragma GCC visibility push(default)
extern "C" int enflow_debug_phase_times(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_phase, sizeof(g_phase));
}
#pragma GCC visibility pop
#else
#define STAMP_E(k, i)
#define STAMP_M(k, i)
#endif

namespace {

constexpr int EPI_THREADS = 512;
constexpr int THREADS = EPI_THREADS + 128; // + one control warpgroup: warp 16 issues the MMAs and TMA copies, 17..19 idle.
// Registers: 20 warps = 5 per scheduler partition -> 96 per thread at launch; the control warpgroup gives most of its
// share back (setmaxnreg.dec) and the four epilogue warpgroups grow to 104 (setmaxnreg.inc); the pool is what the CTA was launched with: 512*104 + 128*56 <= 640*96.
constexpr int EPI_REGS = 104, CTRL_REGS = 56;
constexpr int TE = 64;                     // edges per tile
constexpr int ACT_IMG = 128 * TE * 2;      // one bf16 activation image: 16 KB
constexpr uint32_t TW2_COL = 128, TW3_COL = 256, PARK_COL = 384, TMEM_COLS = 512;   // T_p = 64 p, PARK_p = 384 + 64 p
constexpr int IDX_INTS = 2 * TE + 8;        // row[64], col[64], the four group headers of the tile
constexpr int IDX_BYTES = IDX_INTS * 4;
constexpr int RING = 3;

// Per-tile records written by k_edge_geom_bwd (one fully parallel pass over the edges); `slots` = tiles * 64.
struct GeomView {
    const int* idx;        // [tiles][136]    row[64], col[64], hdr[4]   (TMA -> ring)
    const float* r;        // [slots]         |d|^2
    const float* ds;       // [slots]         d loss / d s  (force branch, egcl.py:71-75)
    const float* d;        // [slots][3]      wrapped coordinate difference
    const float* ddir;     // [slots][3]      direct part of d loss / d d
    const int2* hdr;       // [tiles][4]      per 16-edge group: run id of its first edge (-1: no valid edge), row-start bits
};
__host__ __device__ inline int64_t geom_tiles(int64_t E_cap) { return (E_cap + TE - 1) / TE + 1; }
__host__ __device__ inline GeomView geom_view(unsigned char* base, int64_t E_cap) {
    const int64_t tiles = geom_tiles(E_cap), slots = tiles * TE;
    GeomView g;
    unsigned char* p = base;
    g.idx = reinterpret_cast<const int*>(p); p += tiles * IDX_BYTES;
    g.r = reinterpret_cast<const float*>(p); p += slots * 4;
    g.ds = reinterpret_cast<const float*>(p); p += slots * 4;
    g.d = reinterpret_cast<const float*>(p); p += slots * 12;
    g.ddir = reinterpret_cast<const float*>(p); p += slots * 12;
    g.hdr = reinterpret_cast<const int2*>(p);
    return g;
}

// per-edge geometry and the force-branch seed (enflow/data/base.py:15-19, egcl.py:71-75 differentiated): padding edges of
// the last tile are self-edges of atom 0 with zero seeds
__global__ void __launch_bounds__(256) k_edge_geom_bwd(const int* __restrict__ row, const int* __restrict__ col,
                                                        const int* __restrict__ rowptr, const int* __restrict__ E_dev,
                                                        const float* __restrict__ pos, const float* __restrict__ box,
                                                        const float* __restrict__ s_saved, const float* __restrict__ dF,
                                                        float coords_weight, const int* __restrict__ mis,
                                                        unsigned char* __restrict__ geom, int E_cap) {
    const GeomView g = geom_view(geom, E_cap);
    const int E = E_dev[0] < E_cap ? E_dev[0] : E_cap;
    const int slots = (E + TE - 1) / TE * TE;
    const int lane = threadIdx.x & 31;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < slots; e += gridDim.x * blockDim.x) {      // warp-uniform trip count
        const bool ok = e < E;
        int i = 0, j = 0;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, ds = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
        int start = 0, m = 0;
        if (ok) {
            i = row[e]; j = col[e];
            d0 = wrapf_(pos[(int64_t)i * 3 + 0] - pos[(int64_t)j * 3 + 0], 0.5f * box[(int64_t)i * 3 + 0]);
            d1 = wrapf_(pos[(int64_t)i * 3 + 1] - pos[(int64_t)j * 3 + 1], 0.5f * box[(int64_t)i * 3 + 1]);
            d2 = wrapf_(pos[(int64_t)i * 3 + 2] - pos[(int64_t)j * 3 + 2], 0.5f * box[(int64_t)i * 3 + 2]);
            const int r0 = rowptr[i];
            const int deg = rowptr[i + 1] - r0;
            const float sc = coords_weight / (float)(deg > 1 ? deg : 1);          // helpers.py:70 (Q12)
            const float s = s_saved[e];
            const float dv[3] = {d0, d1, d2};
            float dtr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tr = dv[c] * s;
                const bool pass = (tr >= -100.f) && (tr <= 100.f);               // clamp backward mask
                dtr[c] = pass ? dF[(int64_t)i * 3 + c] * sc : 0.f;
                ds = fmaf(dtr[c], dv[c], ds);
            }
            q0 = dtr[0] * s; q1 = dtr[1] * s; q2 = dtr[2] * s;
            start = e == r0;
            m = mis[i + 1];
        }
        const int tile = e / TE, t = e % TE;
        int* ix = const_cast<int*>(g.idx) + (int64_t)tile * IDX_INTS;
        ix[t] = i; ix[TE + t] = j;
        const_cast<float*>(g.r)[e] = d0 * d0 + d1 * d1 + d2 * d2;
        const_cast<float*>(g.ds)[e] = ds;
        float* gd = const_cast<float*>(g.d) + (int64_t)e * 3;
        gd[0] = d0; gd[1] = d1; gd[2] = d2;
        float* gq = const_cast<float*>(g.ddir) + (int64_t)e * 3;
        gq[0] = q0; gq[1] = q1; gq[2] = q2;
        // header of this edge's 16-group: lanes 0 and 16 of the warp own one each
        const unsigned bits = __ballot_sync(0xffffffffu, start);
        if ((lane & 15) == 0) {
            const int2 h = make_int2(ok ? (e >> 4) + m : -1, (int)((bits >> (lane & 16)) & 0xffffu));
            const_cast<int2*>(g.hdr)[e >> 4] = h;
            reinterpret_cast<int2*>(ix + 2 * TE)[t >> 4] = h;
        }
    }
}

template <bool SPLIT>
struct SmemB {
    static constexpr int NW = SPLIT ? 4 : 2;           // weight images
    static constexpr int NA = SPLIT ? 2 : 1;           // images per activation buffer (hi[, lo])
    static constexpr int NZ = SPLIT ? 1 : 2;           // dz buffers: shared between the pipelines in split mode
    static constexpr size_t ABUF = (size_t)NA * ACT_IMG;
    static constexpr size_t w_off = 0;
    static constexpr size_t x_off = (size_t)NW * tc::IMG_BYTES;
    static constexpr size_t z_off = x_off + 2 * ABUF;
    static constexpr size_t ring_off = z_off + NZ * ABUF;
    static constexpr size_t drp_off = ring_off + RING * IDX_BYTES;      // 3 x 544
    static constexpr size_t bar_off = drp_off + 4 * TE * sizeof(float);
    static constexpr size_t used = bar_off + 128;
    static constexpr size_t total = SPLIT ? 232448 : used + 1024;     // split mode: everything the SM has; base must be 1 KB aligned
};
static_assert(SmemB<true>::used <= 232448, "shared memory budget");

__device__ __forceinline__ uint32_t t_off_(int n, int chunk8) {           // [n][e] image, 128 rows x 64 cols
    return (uint32_t)(n * 128 + ((chunk8 ^ (n & 7)) << 4));
}

template <bool SPLIT>
__device__ __forceinline__ void store8(unsigned char* img, uint32_t off, const float (&x)[8]) {
    if (SPLIT) {
        uint4 hi, lo;
        tc::split2(x[0], x[1], hi.x, lo.x);
        tc::split2(x[2], x[3], hi.y, lo.y);
        tc::split2(x[4], x[5], hi.z, lo.z);
        tc::split2(x[6], x[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(img + off) = hi;
        *reinterpret_cast<uint4*>(img + ACT_IMG + off) = lo;
    } else {
        uint4 hi;
        hi.x = tc::pack_bf16(x[0], x[1]); hi.y = tc::pack_bf16(x[2], x[3]);
        hi.z = tc::pack_bf16(x[4], x[5]); hi.w = tc::pack_bf16(x[6], x[7]);
        *reinterpret_cast<uint4*>(img + off) = hi;
    }
}

// Per-CTA partial layout (floats), identical to the FFMA kernel: dW2 [H*H] | dW3 [H*H] | db2 | db3 | dwc | dwr
constexpr int EDGE_PARTIAL = 2 * ENF_H * ENF_H + 4 * ENF_H;

template <int P> struct Pipe { static constexpr int value = P; };

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_edge_bwd_tc(GeomView gv, const int* __restrict__ E_dev, int E_cap,
              const float* __restrict__ P, const float* __restrict__ S, const float* __restrict__ W1, int e1,
              const float* __restrict__ b2, const float* __restrict__ b3, const float* __restrict__ wc,
              const unsigned char* __restrict__ wimg, const float* __restrict__ dagg,
              float* __restrict__ runs, float* __restrict__ dz1, float* __restrict__ dd_out,
              float* __restrict__ partial, int* __restrict__ status) {
    using L = SmemB<SPLIT>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    if (SPLIT && sm != smem_raw) {                 // no slack in split mode: the window must start 1 KB aligned
        if (threadIdx.x == 0 && status) atomicOr(status, 4);
        return;
    }
    unsigned char* Wimg = sm + L::w_off;
    unsigned char* XB = sm + L::x_off;         // XB + p ABUF: x1^T image [k][e] of pipeline p, then x2^T, then x1^T again
    unsigned char* ZB = sm + L::z_off;         // dz3^T then dz2^T image [n][e]
    int* ring = reinterpret_cast<int*>(sm + L::ring_off);
    float* drp = reinterpret_cast<float*>(sm + L::drp_off);            // [4 quarters][64 edges] partial sums of w_r . dz1
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::bar_off);
    uint64_t* bar_w = bars;                    // weight images have landed
    uint64_t* bar_acc = bars + 1;              // [2] accumulator of pipeline p is ready             (MMA -> epilogue)
    uint64_t* bar_opnd = bars + 3;             // [2] operand images of pipeline p are written, T read (epilogue -> MMA)
    uint64_t* bar_wgd = bars + 5;              // [2] the weight-gradient MMAs of G3 / G4 completed: operands reusable
    uint64_t* bar_idx = bars + 7;              // [3] ring slot filled by TMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        for (int p = 0; p < 2; ++p) {
            tc::mbar_init(bar_acc + p, 1);
            tc::mbar_init(bar_opnd + p, EPI_THREADS / 32);          // one arrival per epilogue warp
            tc::mbar_init(bar_wgd + p, 1);
        }
        for (int s = 0; s < RING; ++s) tc::mbar_init(bar_idx + s, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int E = E_dev[0] < E_cap ? E_dev[0] : E_cap;
    const int tiles = (E + TE - 1) / TE;
    // this CTA's tiles: tile0 .. tile0 + T - 1.  A contiguous range keeps a molecule's rows of P / S / dagg in this SM's L1
    // from tile to tile (measured 3.41 -> 3.38 ms per step on C2 against the strided assignment)
    const int Tc = (tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile0 = blockIdx.x * Tc;
    const int T = tiles - tile0 < 0 ? 0 : (tiles - tile0 < Tc ? tiles - tile0 : Tc);
    const int periods = (T + 2) / 2;                                     // k = 0 .. while 2k - 1 < T

    if (w >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
      if (w == 16) {
        // =========================== MMA / TMA issuing warp ===========================
        const uint32_t w2 = tc::smem_u32(Wimg), w3 = tc::smem_u32(Wimg + (size_t)(SPLIT ? 2 : 1) * tc::IMG_BYTES);
        const uint32_t WLO = tc::IMG_BYTES, ALO = ACT_IMG;
        const uint32_t id_kmn64 = tc::make_idesc(false, true, 64);
        const uint32_t id_mm64 = tc::make_idesc(true, true, 64);
        const uint32_t id_kk128 = tc::make_idesc(false, false, 128);
        // base descriptors (K-major: LBO unused, SBO = 1024; MN-major: LBO = distance between 64-wide M/N blocks)
        const uint64_t dW2k = tc::make_desc(w2, 16, 1024), dW3k = tc::make_desc(w3, 16, 1024);
        const uint64_t dW2m = tc::make_desc(w2, tc::BLK_BYTES, 1024), dW3m = tc::make_desc(w3, tc::BLK_BYTES, 1024);
        uint32_t ph_opnd[2] = {0, 0};
        bool first2 = true, first3 = true;
        auto fetch_idx = [&](int t) {
            if (lane == 0) {
                const int s = t % RING;
                tc::mbar_expect_tx(bar_idx + s, IDX_BYTES);
                tc::bulk_g2s(ring + s * IDX_INTS, gv.idx + (int64_t)(tile0 + t) * IDX_INTS, IDX_BYTES, bar_idx + s);
            }
        };
        // wait for the epilogue warps' arrivals on pipeline p, then the whole warp runs `body` (one elected lane issues)
        int kk = -1, si = 0;                       // (debug stamps only)
        auto step = [&](int p, auto&& body) {
            tc::mbar_wait(bar_opnd + p, ph_opnd[p] & 1);
            ++ph_opnd[p];
            tc::fence_after_sync();
            STAMP_M(kk, 2 * si);
            body(tc::elect_one());
            __syncwarp();
            STAMP_M(kk, 2 * si + 1);
            ++si;
        };
        // k-step strides (bytes): K-major 128-row image: 4 steps of 32 B per 64-column block, blocks 16 KB apart;
        // MN-major: 16 rows = 2 KB per step; K-major [n][e] image with 64 edges: one block, 32 B per step
        auto g_fwd = [&](bool leader, int p, uint64_t dWk) {              // G1 / G2: T_p = W x^T
            const uint64_t dXTm = tc::make_desc(tc::smem_u32(XB + p * L::ABUF), tc::BLK_BYTES, 1024);
            tc::issue_gemm_loop<SPLIT, 2, tc::BLK_BYTES, 32, 8192, 2048>(leader, tmem + 64 * p, dWk, WLO, dXTm, ALO, id_kmn64, false);
            if (leader) tc::mma_commit(bar_acc + p);
        };
        auto g_bwd = [&](bool leader, int p, uint64_t dWm, uint32_t tw_col, bool& first) {      // G3 / G4: dgrad, then wgrad behind it
            const uint32_t zb = tc::smem_u32(ZB + (L::NZ == 2 ? p : 0) * L::ABUF), xb = tc::smem_u32(XB + p * L::ABUF);
            const uint64_t dZm = tc::make_desc(zb, tc::BLK_BYTES, 1024), dZk = tc::make_desc(zb, 16, 1024);
            const uint64_t dXk = tc::make_desc(xb, 16, 1024);
            tc::issue_gemm_loop<SPLIT, 2, 8192, 2048, 8192, 2048>(leader, tmem + 64 * p, dWm, WLO, dZm, ALO, id_mm64, false);
            if (leader) tc::mma_commit(bar_acc + p);             // the epilogue only needs the dgrad ...
            tc::issue_gemm_loop<SPLIT, 1, 0, 32, 0, 32>(leader, tmem + tw_col, dZk, ALO, dXk, ALO, id_kk128, !first);
            if (leader) tc::mma_commit(bar_wgd + p);             // ... operand reuse waits for the wgrad
            first = false;
        };
        if (lane == 0) {
            tc::mbar_expect_tx(bar_w, L::NW * tc::IMG_BYTES);
            for (int i = 0; i < L::NW; ++i) {
                const int src = SPLIT ? i : 2 * i;      // global order: W2_hi, W2_lo, W3_hi, W3_lo
                tc::bulk_g2s(Wimg + (size_t)i * tc::IMG_BYTES, wimg + (size_t)src * tc::IMG_BYTES, tc::IMG_BYTES, bar_w);
            }
        }
        // ring slot t % 3 is free for tile t as soon as every epilogue warp has passed E3(t - 3) (the last reader of that
        // slot): the first three tiles are requested here, tile t + 3 together with G4(t)
        for (int t = 0; t < RING && t < T; ++t) fetch_idx(t);
        tc::mbar_wait(bar_w, 0);
        if (T > 0) {
            step(0, [&](bool ld) { g_fwd(ld, 0, dW2k); });                  // X(0)  -> G1(0)
            step(0, [&](bool ld) { g_fwd(ld, 0, dW3k); });                  // E1(0) -> G2(0)
        }
        for (int k = 0; k < periods; ++k) {
            const int a = 2 * k, b = a + 1, a2 = a + 2;
            kk = k; si = 0;
            if (a < T) step(0, [&](bool ld) { g_bwd(ld, 0, dW3m, TW3_COL, first3); });          // E2(a)  -> G3(a)
            if (b < T) step(1, [&](bool ld) { g_fwd(ld, 1, dW2k); });                           // X(b)   -> G1(b)
            if (a < T) step(0, [&](bool ld) { if (a + RING < T) fetch_idx(a + RING); g_bwd(ld, 0, dW2m, TW2_COL, first2); });   // E3(a)  -> G4(a)
            if (b < T) step(1, [&](bool ld) { g_fwd(ld, 1, dW3k); });                           // E1(b)  -> G2(b)
            if (a2 < T) step(0, [&](bool ld) { g_fwd(ld, 0, dW2k); });                          // X(a')  -> G1(a')
            if (b < T) step(1, [&](bool ld) { g_bwd(ld, 1, dW3m, TW3_COL, first3); });          // E2(b)  -> G3(b)
            if (a2 < T) step(0, [&](bool ld) { g_fwd(ld, 0, dW3k); });                          // E1(a') -> G2(a')
            if (b < T) step(1, [&](bool ld) { if (b + RING < T) fetch_idx(b + RING); g_bwd(ld, 1, dW2m, TW2_COL, first2); });   // E3(b)  -> G4(b)
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
        // =========================== epilogue warps ===========================
        const int q = w & 3, cg = w >> 2;
        const int n = 32 * q + lane;           // hidden unit == TMEM lane
        const int ec = 16 * cg;                // first tile edge of this thread's 16 columns
        const float b2n = b2[n], b3n = b3[n], wcn = wc[n], wrn = W1[n * e1 + e1 - 1];
        const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
        float gb2 = 0.f, gb3 = 0.f, gwc = 0.f, gwr = 0.f;
        float sa[16], sb[16];                  // per-pipeline state: z1 of the pipeline's tile, from its gather to E3

        int kq = -1, tq = 0;                    // (debug stamps only)
        // mbarrier arrivals go through the same L1/shared-memory data pipe as the MMA operand reads, and 32 lanes arriving
        // on one barrier word are 32 serialised accesses (ncu: the per-thread form was 15 % of that pipe): one lane per
        // warp arrives, __syncwarp orders the other lanes' writes before it
        auto wait_bar = [&](uint64_t* bar, uint32_t parity) { tc::mbar_wait(bar, parity); };
        auto wait_acc = [&](int p, uint32_t parity) {
            STAMP_E(kq, 3 * tq);
            wait_bar(bar_acc + p, parity);
            tc::fence_after_sync();
            STAMP_E(kq, 3 * tq + 1);
        };
        auto arrive = [&](int p) {             // this warp's operand writes and TMEM reads of the task are done
            tc::fence_before_sync();
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_opnd + p);
            STAMP_E(kq, 3 * tq + 2);
        };
        // z1 = P[row] + S[col] + w_r r for this thread's (hidden unit, 16 edges); row reads coalesce over the 32 hidden
        // units of a warp
        // The P row (and in E3 the dagg row) is requested only where the row changes inside the thread's 16 edges (row-start
        // bits of the group header): the gathers are bound by the number of outstanding 128-byte requests per SM, and CSR
        // order makes most of the row requests repeats.
        auto gather_z1 = [&](int t, float (&z)[16]) {
            const int s = t % RING;
            wait_bar(bar_idx + s, (uint32_t)(t / RING) & 1);
            const int* ir = ring + s * IDX_INTS + ec;
            const int bits = ring[s * IDX_INTS + 2 * TE + 2 * cg + 1];
            const float4* r4 = reinterpret_cast<const float4*>(gv.r + (int64_t)(tile0 + t) * TE + ec);
            // requests first (every load has its own destination register), the repeats are filled in afterwards
            float pz[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int4 ri = *reinterpret_cast<const int4*>(ir + 4 * c);
                const int4 ci = *reinterpret_cast<const int4*>(ir + TE + 4 * c);
                const int rw[4] = {ri.x, ri.y, ri.z, ri.w}, cl[4] = {ci.x, ci.y, ci.z, ci.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    pz[j] = 0.f;
                    if (j == 0 || ((bits >> j) & 1)) pz[j] = __ldg(P + (int64_t)rw[u] * ENF_H + n);
                    z[j] = __ldg(S + (int64_t)cl[u] * ENF_H + n);
                }
            }
            float pv = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 rr = __ldg(r4 + c);
                const float rq[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    pv = (j == 0 || ((bits >> j) & 1)) ? pz[j] : pv;
                    z[j] = fmaf(wrn, rq[u], pv + z[j]);
                }
            }
        };
        // x1^T = silu(z1)^T into the [hidden][edge] image; DS: z1 is replaced by silu'(z1)
        auto put_x1 = [&](unsigned char* X, float (&z1)[16], auto DS) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float z = z1[8 * ch + j];
                    const float sg = tc::sigmoid_t<SPLIT>(z);
                    x[j] = z * sg;
                    if (decltype(DS)::value) z1[8 * ch + j] = fmaf(x[j], 1.0f - sg, sg);       // silu'(z) = s + z s (1 - s)
                }
                store8<SPLIT>(X, t_off_(n, 2 * cg + ch), x);
            }
        };
        // The gather of a pipeline's next tile is issued one to two tasks before its x1 image is written, in front of a
        // wait that is long anyway (the state registers of that pipeline are free from E3 on: silu'(z1) is parked in TMEM)
        auto prefetch_z1 = [&](int t, float (&z)[16]) {
            if (t >= 0 && t < T) gather_z1(t, z);
        };

        // ---- X(t): gather + x1 image of tile t (first tiles only; later ones ride on E4 of tile t - 2)
        auto taskX = [&](auto PP, int t, float (&st)[16]) {
            constexpr int p = decltype(PP)::value;
            gather_z1(t, st);
            put_x1(XB + p * L::ABUF, st, Pipe<0>{});
            arrive(p);
        };
        // ---- E1(t): z2 -> x2^T image over x1^T, silu'(z2) parked in TMEM
        auto taskE1 = [&](auto PP, int tg, float (&sg_other)[16]) {
            constexpr int p = decltype(PP)::value;
            prefetch_z1(tg, sg_other);              // next tile of the OTHER pipeline (G1 of this one is usually still running)
            wait_acc(p, 0);
            float v[16];
            tc::tmem_ld16(lane_base + 64 * p + ec, v);
            unsigned char* X = XB + p * L::ABUF;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int jj = 8 * ch + j;
                    const float z = v[jj] + b2n;
                    const float sg = tc::sigmoid_t<SPLIT>(z);
                    x[j] = z * sg;
                    v[jj] = fmaf(x[j], 1.0f - sg, sg);
                }
                store8<SPLIT>(X, t_off_(n, 2 * cg + ch), x);
            }
            tc::tmem_st16(lane_base + PARK_COL + 64 * p + ec, v);
            arrive(p);
        };
        // ---- E2(t): z3 -> dz3^T image; t_prev_other >= 0: tile of the other pipeline whose G4 still reads the shared buffer
        auto taskE2 = [&](auto PP, int t, int tg, float (&sg_other)[16]) {
            constexpr int p = decltype(PP)::value;
            const float4* d4 = reinterpret_cast<const float4*>(gv.ds + (int64_t)(tile0 + t) * TE + ec);
            float dsv[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 x = __ldg(d4 + c);
                dsv[4 * c] = x.x; dsv[4 * c + 1] = x.y; dsv[4 * c + 2] = x.z; dsv[4 * c + 3] = x.w;
            }
            wait_acc(p, 1);
            float v[16];
            tc::tmem_ld16(lane_base + 64 * p + ec, v);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float z = v[jj] + b3n;
                const float sg = tc::sigmoid_t<SPLIT>(z);
                const float ds = dsv[jj];
                const float x3 = z * sg;
                gwc = fmaf(ds, x3, gwc);
                const float dz = ds * wcn * fmaf(x3, 1.0f - sg, sg);
                gb3 += dz;
                v[jj] = dz;
            }
            prefetch_z1(tg, sg_other);              // next tile of the OTHER pipeline, in front of the one exposed MMA wait
            if (L::NZ == 1 && t >= 1) wait_bar(bar_wgd + (1 - p), 1);       // G4(t - 1) has released the shared dz buffer
            unsigned char* Z = ZB + (L::NZ == 2 ? p : 0) * L::ABUF;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[8 * ch + j];
                store8<SPLIT>(Z, t_off_(n, 2 * cg + ch), x);          // dz3^T image [n][e]
            }
            arrive(p);
        };
        // ---- E3(t): dx2 -> dz2^T image, x1^T image rebuilt (x2^T is dead), silu'(z1) kept
        auto taskE3 = [&](auto PP, int t, float (&st)[16]) {
            constexpr int p = decltype(PP)::value;
            const int e0 = (tile0 + t) * TE;
            const int* ir = ring + (t % RING) * IDX_INTS + ec;
            const int bits = ring[(t % RING) * IDX_INTS + 2 * TE + 2 * cg + 1];
            float pf[16];                               // dagg rows: requested where the row changes, repeats filled in below
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int4 ri = *reinterpret_cast<const int4*>(ir + 4 * c);
                const int rw[4] = {ri.x, ri.y, ri.z, ri.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = 4 * c + u;
                    pf[j] = 0.f;
                    if (j == 0 || ((bits >> j) & 1)) pf[j] = __ldg(dagg + (int64_t)rw[u] * ENF_H + n);
                }
            }
            wait_acc(p, 0);
            float v[16];
            tc::tmem_ld16(lane_base + 64 * p + ec, v);
            {
                float dsl2[16];
                tc::tmem_ld16(lane_base + PARK_COL + 64 * p + ec, dsl2);
                float av = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    av = (j == 0 || ((bits >> j) & 1)) ? pf[j] : av;             // repeats of the row's dagg
                    v[j] = (v[j] + ((e0 + ec + j < E) ? av : 0.f)) * dsl2[j];
                    gb2 += v[j];
                }
            }
            wait_bar(bar_wgd + p, 0);                    // dz3^T / x2^T were still being read by the TW3 MMAs until here
            unsigned char* Z = ZB + (L::NZ == 2 ? p : 0) * L::ABUF;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[8 * ch + j];
                store8<SPLIT>(Z, t_off_(n, 2 * cg + ch), x);          // dz2^T image [n][e]
            }
            put_x1(XB + p * L::ABUF, st, Pipe<1>{});
            tc::tmem_st16(lane_base + PARK_COL + 64 * p + ec, st);      // silu'(z1) for E4 (silu'(z2) has been consumed)
            arrive(p);
        };
        // ---- E4(t) + X(t2 = t + 2): dx1 -> dz1, per-run row sums (dP), d loss / d r; then the next tile of this pipeline
        auto taskE4X = [&](auto PP, int t, int t2, float (&st)[16]) {
            constexpr int p = decltype(PP)::value;
            const bool has_t = t >= 0 && t < T, has_x = t2 < T;
            if (has_t) {
                const int tile = tile0 + t;
                const int e0 = tile * TE;
                const int2 hdr = __ldg(gv.hdr + (int64_t)tile * 4 + cg);
                const float4* r4 = reinterpret_cast<const float4*>(gv.r + (int64_t)e0 + ec);
                float rv[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 x = __ldg(r4 + c);
                    rv[4 * c] = x.x; rv[4 * c + 1] = x.y; rv[4 * c + 2] = x.z; rv[4 * c + 3] = x.w;
                }
                // d loss / d d of edge (ec + 4 q + lane) is finished by lane < 4 of this warp: its inputs are requested now
                const int md = ec + 4 * q + (lane & 3);
                float dv[3], qv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    dv[c] = __ldg(gv.d + (int64_t)(e0 + md) * 3 + c);
                    qv[c] = __ldg(gv.ddir + (int64_t)(e0 + md) * 3 + c);
                }
                wait_acc(p, 1);
                float v[16];
                tc::tmem_ld16(lane_base + 64 * p + ec, v);
                float ds1[16];
                tc::tmem_ld16(lane_base + PARK_COL + 64 * p + ec, ds1);
                // dz1 = dx1 * silu'(z1): stored per edge (for the column-grouped sum dS) and reduced over each row's
                // edges into per-run partials (dP, see segment.cu; tc::run_sums16)
                const int nvalid = E - (e0 + ec);                     // edges [0, nvalid) of this thread's 16 exist
                float* dzp = dz1 + (int64_t)(e0 + ec) * ENF_H + n;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] *= ds1[j];
                    gwr = fmaf(v[j], rv[j], gwr);
                }
                if (nvalid >= 16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dzp[(int64_t)j * ENF_H] = v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < nvalid) dzp[(int64_t)j * ENF_H] = v[j];
                }
                tc::run_sums16(v, hdr.x, (unsigned)hdr.y, nvalid, runs, n);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= wrn;
                const float tsum = tc::warp_transpose_sum16(v, lane);
                if (!(lane & 1)) drp[q * TE + ec + (lane >> 1)] = tsum;
                tc::named_bar_sync(1 + cg, 128);                     // the four quarter-warps of this edge group
                if (lane < 4 && e0 + md < E) {
                    const float dr2 = 2.0f * ((drp[md] + drp[TE + md]) + (drp[2 * TE + md] + drp[3 * TE + md]));
#pragma unroll
                    for (int c = 0; c < 3; ++c) dd_out[(int64_t)(e0 + md) * 3 + c] = fmaf(dr2, dv[c], qv[c]);
                }
            }
            if (has_x) {                                             // st holds z1 of tile t2 (prefetch_z1)
                if (has_t) wait_bar(bar_wgd + p, 1);                 // TW2 += dz2^T x1 of tile t still reads the x buffer
                put_x1(XB + p * L::ABUF, st, Pipe<0>{});
                arrive(p);
            }
        };

        if (T > 0) {
            taskX(Pipe<0>{}, 0, sa);
            taskE1(Pipe<0>{}, -1, sb);
        }
        for (int k = 0; k < periods; ++k) {
            const int a = 2 * k, b = a + 1, a2 = a + 2;
            kq = k;
            tq = 0; if (a < T) taskE2(Pipe<0>{}, a, b, sb);          // + gather of tile b (x1 image in the next task)
            tq = 1; taskE4X(Pipe<1>{}, b - 2, b, sb);
            tq = 2; if (a < T) taskE3(Pipe<0>{}, a, sa);
            tq = 3; if (b < T) taskE1(Pipe<1>{}, a2, sa);            // + gather of tile a' (x1 image in the next task)
            tq = 4; taskE4X(Pipe<0>{}, a, a2, sa);
            tq = 5; if (b < T) taskE2(Pipe<1>{}, b, -1, sa);
            tq = 6; if (a2 < T) taskE1(Pipe<0>{}, -1, sb);
            tq = 7; if (b < T) taskE3(Pipe<1>{}, b, sb);
            STAMP_E(kq, 24);
        }
        // every MMA of this CTA has completed before TW2 / TW3 are read
        if (T > 0) wait_bar(bar_wgd + 0, 1);
        if (T > 1) wait_bar(bar_wgd + 1, 1);
        tc::fence_after_sync();
        // ---- per-CTA partials: weight gradients from TMEM
        float* my = partial + (int64_t)blockIdx.x * EDGE_PARTIAL;
        {
            float v[32];
#pragma unroll 1
            for (int mat = 0; mat < 2; ++mat) {
                if (T == 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;        // this CTA had no tile: TMEM was never written
                } else {
                    tc::tmem_ld32(lane_base + (mat ? TW3_COL : TW2_COL) + 32 * cg, v);
                }
                float4* dst = reinterpret_cast<float4*>(my + mat * ENF_H * ENF_H + n * ENF_H + 32 * cg);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        float* red = reinterpret_cast<float*>(XB);          // [4 kinds][4 cg][128]; every MMA that read XB is complete
        red[(0 * 4 + cg) * ENF_H + n] = gb2;
        red[(1 * 4 + cg) * ENF_H + n] = gb3;
        red[(2 * 4 + cg) * ENF_H + n] = gwc;
        red[(3 * 4 + cg) * ENF_H + n] = gwr;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < 4 * ENF_H) {
        const float* red = reinterpret_cast<const float*>(XB);
        float* my = partial + (int64_t)blockIdx.x * EDGE_PARTIAL;
        const int a = tid / ENF_H, k = tid % ENF_H;
        my[2 * ENF_H * ENF_H + tid] = (red[(a * 4 + 0) * ENF_H + k] + red[(a * 4 + 1) * ENF_H + k]) +
                                      (red[(a * 4 + 2) * ENF_H + k] + red[(a * 4 + 3) * ENF_H + k]);
    }
    if (w == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

int enf_edge_reduce_partials(const float* partial, int n_cta, float* lgrad, int nf, cudaStream_t st);

// bytes of the per-tile records k_edge_geom_bwd writes for a capacity of E_cap edges (16-byte aligned buffer)
int64_t enf_edge_bwd_geom_bytes(int E_cap) {
    const int64_t tiles = geom_tiles(E_cap);
    return tiles * IDX_BYTES + tiles * TE * (4 + 4 + 12 + 12) + tiles * 4 * (int64_t)sizeof(int2) + 256;
}

// the per-tile edge records of the backward kernel (needs dF of this layer's coupling step; nothing else of the layer)
int enf_edge_bwd_tc_geom(const int* row, const int* col, const int* rowptr, const int* E_dev, int E_cap, const float* pos,
                         const float* box, const float* s_saved, const float* dF, float coords_weight, const int* mis,
                         unsigned char* geom, cudaStream_t st) {
    if (E_cap == 0) return ENF_OK;
    const int slots = (E_cap + TE - 1) / TE * TE;
    int ggrid = (slots + 255) / 256;
    if (ggrid > 8 * enf_num_sms()) ggrid = 8 * enf_num_sms();
    enf_time_begin(TK_EDGE_GEOM, st);
    enf_count_launch(), k_edge_geom_bwd<<<ggrid, 256, 0, st>>>(row, col, rowptr, E_dev, pos, box, s_saved, dF, coords_weight, mis, geom, E_cap);
    enf_time_end(st);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// the backward kernel on `st` (the records must be complete on `st`), the reduction of its per-CTA weight-gradient
// partials on `st_red` (== st: in line)
int enf_edge_bwd_tc(int mode, const int* E_dev, int E_cap, const float* P, const float* S, const float* lp,
                    const unsigned char* wimg, int nf, const float* dagg, float* runs, float* dz1, float* dd, float* lgrad,
                    float* partial, unsigned char* geom, int* status, cudaStream_t st, cudaStream_t st_red) {
    if (E_cap == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const int grid = enf_num_sms();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_edge_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemB<true>::total);
        cudaFuncSetAttribute(k_edge_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemB<false>::total);
        attr = true;
    }
    const GeomView gv = geom_view(geom, E_cap);
    enf_time_begin(TK_EDGE_BWD, st);
    if (mode == 1)
        enf_count_launch(), k_edge_bwd_tc<true><<<grid, THREADS, SmemB<true>::total, st>>>(
            gv, E_dev, E_cap, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], wimg, dagg,
            runs, dz1, dd, partial, status);
    else
        enf_count_launch(), k_edge_bwd_tc<false><<<grid, THREADS, SmemB<false>::total, st>>>(
            gv, E_dev, E_cap, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], wimg, dagg,
            runs, dz1, dd, partial, status);
    enf_time_end(st);
    ENF_CHECK_LAUNCH();
    enf_chain(st, st_red);
    enf_time_begin(TK_EDGE_REDUCE, st_red);
    const int rc = enf_edge_reduce_partials(partial, grid, lgrad, nf, st_red);
    enf_time_end(st_red);
    return rc;
}
