// Implicit neighbour lists for the fully connected regime ("or implicit all-pairs when provably in
// the FC regime"; reference semantics enflow/data/base.py:122-144 + utils/helpers.py:15-29, quirks Q6-Q11).
//
// When a molecule satisfies the four conditions below, the reference's Data.edges returns exactly all ordered pairs
// (i, j), i != j, of the molecule, sorted by row then column.  The list then follows from index arithmetic and is the
// same at every coupling step, so it is built ONCE per pass (k_fc_offsets + k_fc_fill: CSR rows, the column-grouped
// permutation for the backward scatter, edge count) and each coupling step only re-proves the regime on its own
// positions (k_fc_check, one launch) instead of running K0 (survivors, two hit passes, three scans) and, in the
// backward pass, the column permutation.  A violation raises status bit 8; the caller repeats the pass with K0.
//
// Conditions, per molecule with box b = box[first atom] and cut-off rc (fp32, base.py:171):
//   (a) every pair of atoms passes K0's own test: the fp64 distance square, evaluated with K0's operation order, is
//       < the fp32 product rc*rc (Q9);
//   (b) no shifted image can produce a hit: b_d - (max_d - min_d) > rc (1 + 1e-6) in every dimension (a shifted
//       copy is at least that far from every atom along a shifted axis; conservative);
//   (c) the ellipsoid pre-filter (Q7), evaluated with K0's operations, keeps either all atoms of an image or none of
//       them, for each of the 27 images: the compacted survivor list then starts with a complete image in atom order,
//       so the reference's second remap through id_mapping (Q6) is the identity on atom ids;
//   (d) the unshifted image (the last one, Q10) survives completely.
// (a)+(b): the hit list is exactly (unshifted atom i, atom j) for all i, j; (c)+(d): labels are the atom ids, self
// pairs drop out (Q11), order is (i, j) row-major.  tests/test_gpu_fc.py holds the result bit-exact against K0.
#include "common.cuh"

namespace {

__device__ __forceinline__ double shift_of(int idx, double L) {   // helpers.py:17: [-L, +L, 0]
    return idx == 0 ? -L : (idx == 1 ? L : 0.0);
}

constexpr int FC_T = 128, FC_SM = 512;

__global__ void __launch_bounds__(FC_T) k_fc_check(const float* __restrict__ pos, const float* __restrict__ box,
                                                    const float* __restrict__ r_cut, const int* __restrict__ mol_off,
                                                    int* __restrict__ status) {
    __shared__ double sx[FC_SM], sy[FC_SM], sz[FC_SM];
    __shared__ int img_in[27], bad;
    __shared__ float lo[3][FC_T / 32], hi[3][FC_T / 32];
    const int m = blockIdx.x;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    if (n <= 1) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (n > FC_SM) {                         // larger molecules stay on K0
        if (tid == 0) atomicOr(status, 8);
        return;
    }
    const float rcf = r_cut[m];
    const double rc = (double)rcf, r_sq = (double)__fmul_rn(rcf, rcf);
    const double bx = (double)box[(int64_t)o * 3 + 0], by = (double)box[(int64_t)o * 3 + 1], bz = (double)box[(int64_t)o * 3 + 2];
    if (tid < 27) img_in[tid] = 0;
    if (tid == 0) bad = 0;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int a = tid; a < n; a += FC_T) {
        const float px = pos[(int64_t)(o + a) * 3 + 0], py = pos[(int64_t)(o + a) * 3 + 1], pz = pos[(int64_t)(o + a) * 3 + 2];
        sx[a] = (double)px; sy[a] = (double)py; sz[a] = (double)pz;
        mn[0] = fminf(mn[0], px); mx[0] = fmaxf(mx[0], px);
        mn[1] = fminf(mn[1], py); mx[1] = fmaxf(mx[1], py);
        mn[2] = fminf(mn[2], pz); mx[2] = fmaxf(mx[2], pz);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], s));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], s));
        }
        if (lane == 0) { lo[c][wid] = mn[c]; hi[c][wid] = mx[c]; }
    }
    __syncthreads();
    int fail = 0;
    // (b) margins
    if (tid < 3) {
        float a = lo[tid][0], b = hi[tid][0];
        for (int w = 1; w < FC_T / 32; ++w) { a = fminf(a, lo[tid][w]); b = fmaxf(b, hi[tid][w]); }
        const double bd = tid == 0 ? bx : (tid == 1 ? by : bz);
        if (!(bd - ((double)b - (double)a) > rc * (1.0 + 1e-6))) fail = 1;
    }
    // (c) ellipsoid pre-filter with K0's operations (edges.cu k_edges_survivors)
    const double ex = __dadd_rn(bx, rc), ey = __dadd_rn(by, rc), ez = __dadd_rn(bz, rc);
    for (int ip = tid; ip < 27 * n; ip += FC_T) {
        const int k = ip / n, a = ip - k * n;
        const double px = __dadd_rn(sx[a], shift_of(k % 3, bx));
        const double py = __dadd_rn(sy[a], shift_of((k / 3) % 3, by));
        const double pz = __dadd_rn(sz[a], shift_of(k / 9, bz));
        const double qx = __ddiv_rn(px, ex), qy = __ddiv_rn(py, ey), qz = __ddiv_rn(pz, ez);
        const double q = __dadd_rn(__dadd_rn(__dmul_rn(qx, qx), __dmul_rn(qy, qy)), __dmul_rn(qz, qz));
        if (q <= 1.0) atomicAdd(&img_in[k], 1);
    }
    // (a) all pairs with K0's distance arithmetic (edges.cu k_edges_hits); the square is symmetric in the pair
    const int pairs = n * (n - 1) / 2;
    for (int t = tid; t < pairs; t += FC_T) {
        // t -> (i, j), i < j: row i holds n - 1 - i pairs
        int i = (int)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * t)) * 0.5);
        while (i > 0 && i * (2 * n - i - 1) / 2 > t) --i;
        while ((i + 1) * (2 * n - i - 2) / 2 <= t) ++i;
        const int j = i + 1 + (t - i * (2 * n - i - 1) / 2);
        const double dx = __dsub_rn(sx[i], sx[j]), dy = __dsub_rn(sy[i], sy[j]), dz = __dsub_rn(sz[i], sz[j]);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!(d2 < r_sq)) fail = 1;
    }
    __syncthreads();
    if (tid < 27 && img_in[tid] != 0 && img_in[tid] != n) fail = 1;      // (c)
    if (tid == 26 && img_in[26] != n) fail = 1;                           // (d)
    if (fail) atomicOr(&bad, 1);
    __syncthreads();
    if (tid == 0 && bad) atomicOr(status, 8);
}

// eoff[m] = sum over molecules before m of n (n - 1); single CTA
__global__ void __launch_bounds__(1024) k_fc_offsets(const int* __restrict__ mol_off, int B, int E_cap,
                                                      int* __restrict__ eoff, int* __restrict__ E_dev,
                                                      int* __restrict__ status) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int m = base + tid;
        long long v = 0;
        if (m < B) {
            const long long n = mol_off[m + 1] - mol_off[m];
            v = n * (n - 1);
        }
        long long x = v;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, s);
            if (lane >= s) x += y;
        }
        if (lane == 31) wsum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            long long w = wsum[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, s);
                if (lane >= s) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (wid ? wsum[wid - 1] : 0) + x - v;
        if (m < B) eoff[m] = (int)(before < 0x7fffffffLL ? before : 0x7fffffffLL);
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) {
        const long long E = carry;
        eoff[B] = (int)(E < 0x7fffffffLL ? E : 0x7fffffffLL);
        E_dev[0] = (int)(E < E_cap ? E : E_cap);
        E_dev[1] = (int)(E < 0x7fffffffLL ? E : 0x7fffffffLL);
        if (E > E_cap) atomicOr(status, 1);
    }
}

// CSR rows, column-grouped permutation (stable: ascending row inside a column) of the all-pairs list
__global__ void __launch_bounds__(256) k_fc_fill(const int* __restrict__ mol_off, const int* __restrict__ eoff, int B,
                                                  int N, int E_cap, int* __restrict__ row, int* __restrict__ col,
                                                  int* __restrict__ rowptr, int* __restrict__ colptr,
                                                  int* __restrict__ perm) {
    const int m = blockIdx.x;
    const int o = mol_off[m], n = mol_off[m + 1] - o;
    const int eb = eoff[m];
    const int deg = n - 1;
    const int step = blockDim.x * gridDim.y, t0 = threadIdx.x + blockDim.x * blockIdx.y;
    for (int a = t0; a < n; a += step) {
        rowptr[o + a] = eb + a * deg;
        if (colptr) colptr[o + a] = eb + a * deg;
    }
    if (m == B - 1 && t0 == 0) {
        rowptr[N] = eoff[B];
        if (colptr) colptr[N] = eoff[B];
    }
    const int ne = n * deg;
    for (int t = t0; t < ne; t += step) {
        const int i = t / deg, jj = t - i * deg;
        const int e = eb + t;
        if (e < E_cap) {
            row[e] = o + i;
            col[e] = o + jj + (jj >= i);
            // position t of the column-grouped order: column i (same index arithmetic), k-th source row
            if (perm) {
                const int src_row = jj + (jj >= i);                   // the row that holds column i as its edge ...
                perm[e] = eb + src_row * deg + (i - (i > src_row));   // ... at this offset inside that row
            }
        }
    }
}

}  // namespace

// status |= 8 when some molecule is not provably in the fully connected regime at these positions
int enf_fc_check(const float* pos, const float* box, const float* r_cut, const int* mol_off, int B, int* status,
                 cudaStream_t st) {
    if (B == 0) return ENF_OK;
    enf_count_launch(), k_fc_check<<<B, FC_T, 0, st>>>(pos, box, r_cut, mol_off, status);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// the all-pairs list of every molecule: row/col [E_cap], rowptr [N+1], E_dev [2], optional colptr [N+1] / perm [E_cap];
// eoff: B + 1 ints of scratch
int enf_fc_build(const int* mol_off, int B, int N, int E_cap, int* row, int* col, int* rowptr, int* E_dev, int* colptr,
                 int* perm, int* eoff, int* status, cudaStream_t st) {
    if (B == 0 || N == 0) {
        cudaMemsetAsync(rowptr, 0, sizeof(int) * ((size_t)N + 1), st);
        cudaMemsetAsync(E_dev, 0, sizeof(int) * 2, st);
        return ENF_OK;
    }
    enf_count_launch(), k_fc_offsets<<<1, 1024, 0, st>>>(mol_off, B, E_cap, eoff, E_dev, status);
    int ysplit = (int)((int64_t)N / B / 64);
    ysplit = ysplit < 1 ? 1 : (ysplit > 32 ? 32 : ysplit);
    enf_count_launch(), k_fc_fill<<<dim3(B, ysplit), 256, 0, st>>>(mol_off, eoff, B, N, E_cap, row, col, rowptr, colptr, perm);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
