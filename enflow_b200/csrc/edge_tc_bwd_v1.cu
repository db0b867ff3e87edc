// K1 backward on tcgen05: recompute + dgrad + wgrad of the two dense edge layers, per 64-edge tile,
// without reading any saved [E,H] activation (enflow/nn/egcl.py:57-63,71-75 differentiated by hand).
//
// All GEMMs keep the accumulator TRANSPOSED (TMEM lane = hidden unit, column = edge or weight column):
//   T1 [n][e] = W2 x1^T            A = W2 image (K-major)          B = x1^T image [k][e] (MN-major)
//   T2 [n][e] = W3 x2^T            A = W3 image (K-major)          B = x2^T image [k][e] (MN-major)
//   TW3[n][k] += dz3^T x2          A = dz3^T image [n][e] (K-major) B = x2^T image [k][e] (K-major)
//   T2 [k][e] = W3^T dz3^T         A = W3 image (MN-major)         B = dz3^T image [n][e] (MN-major)
//   TW2[n][k] += dz2^T x1          A = dz2^T image [n][e] (K-major) B = x1^T image [k][e] (K-major)
//   T1 [k][e] = W2^T dz2^T         A = W2 image (MN-major)         B = dz2^T image [n][e] (MN-major)
// One swizzled image per operand serves every view (tc_common.cuh).  The weight-gradient accumulators TW2/TW3
// stay in TMEM for the whole life of the CTA and are written once at the end as a per-CTA partial; partials are
// combined in CTA order (deterministic).  x1 is regenerated rather than kept: with the bf16x3 operand split
// (hi/lo images) shared memory holds the two weight matrices (128 KB) plus two activation buffers (64 KB).
//
// Thread map: 16 warps; warp w owns TMEM lanes [32 (w%4), +32) = hidden units n and tile edges [16 (w/4), +16).
// Tiles are software-pipelined per CTA: per-tile edge records (k_edge_geom_bwd) arrive by TMA two tiles ahead, the z1
// gather and the first MMA of tile t+1 are issued during tile t (profiles/r1c_phase_times.txt).
#include "common.cuh"
#include "tc_common.cuh"

// Debug builds (make PHASE=1): clock64 stamps at the phase boundaries of four tiles of one CTA, read back by
// tools/phase_times.py.  Not part of the product library (the extra symbol is not in include/enflow_b200.h).
#ifdef ENF_PHASE_TIMING
__device__ long long g_phase[4 * 16];
#define STAMP(k) do { if (blockIdx.x == 3 && tid == 37 && tile_no >= 8 && tile_no < 12) g_phase[(tile_no - 8) * 16 + (k)] = clock64(); } while (0)
#pragma GCC visibility push(default)
extern "C" int enflow_debug_phase_times(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_phase, sizeof(long long) * 64);
}
#pragma GCC visibility pop
#else
#define STAMP(k)
#endif

namespace {

constexpr int THREADS = 512;
constexpr int TE = 64;                     // edges per tile
constexpr int ACT_IMG = 128 * TE * 2;      // one bf16 activation image: 16 KB
constexpr uint32_t T1_COL = 0, T2_COL = 64, TW2_COL = 128, TW3_COL = 256, T1B_COL = 384, TMEM_COLS = 512;   // T1 alternates per tile

// Per-tile record.  The first GEOM_BYTES are produced for every tile by k_edge_geom_bwd (one fully parallel pass over
// the edges) and arrive in shared memory by one TMA bulk copy two tiles ahead; dr_part is kernel-local scratch.
struct TileInfoB {
    int row[TE], col[TE], valid[TE], start[TE], mis[TE];
    float d[TE][3], r[TE], ds[TE], ddir[TE][3];
    float dr_part[4][TE];
};
constexpr int GEOM_BYTES = 13 * TE * 4;
static_assert(offsetof(TileInfoB, dr_part) == GEOM_BYTES && GEOM_BYTES % 16 == 0 && sizeof(TileInfoB) % 16 == 0, "tile record layout");

// per-edge geometry and the force-branch seed (enflow/data/base.py:15-19, egcl.py:71-75 differentiated): padding edges of
// the last tile are self-edges of atom 0 with zero seeds
__global__ void __launch_bounds__(256) k_edge_geom_bwd(const int* __restrict__ row, const int* __restrict__ col,
                                                        const int* __restrict__ rowptr, const int* __restrict__ E_dev,
                                                        const float* __restrict__ pos, const float* __restrict__ box,
                                                        const float* __restrict__ s_saved, const float* __restrict__ dF,
                                                        float coords_weight, const int* __restrict__ mis,
                                                        unsigned char* __restrict__ geom) {
    const int E = E_dev[0];
    const int slots = (E + TE - 1) / TE * TE;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < slots; e += gridDim.x * blockDim.x) {
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
        TileInfoB& ti = *reinterpret_cast<TileInfoB*>(geom + (int64_t)(e / TE) * GEOM_BYTES);     // only the record part exists
        const int t = e % TE;
        ti.row[t] = i; ti.col[t] = j; ti.valid[t] = ok; ti.start[t] = start; ti.mis[t] = m;
        ti.d[t][0] = d0; ti.d[t][1] = d1; ti.d[t][2] = d2;
        ti.r[t] = d0 * d0 + d1 * d1 + d2 * d2;
        ti.ds[t] = ds;
        ti.ddir[t][0] = q0; ti.ddir[t][1] = q1; ti.ddir[t][2] = q2;
    }
}

template <bool SPLIT>
struct SmemB {
    static constexpr int NW = SPLIT ? 4 : 2;
    static constexpr int NA = SPLIT ? 2 : 1;
    static constexpr size_t w_off = 0;
    static constexpr size_t x_off = (size_t)NW * tc::IMG_BYTES;
    static constexpr size_t z_off = x_off + (size_t)NA * ACT_IMG;
    static constexpr size_t t_off = z_off + (size_t)NA * ACT_IMG;
    static constexpr size_t bar_off = (t_off + 3 * sizeof(TileInfoB) + 15) / 16 * 16;      // tile info: three tiles in flight
    static constexpr size_t total = bar_off + 64 + 1024;
};

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

// sum over the 32 lanes of 16 per-lane values; lane l receives column (l & 15)
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Per-CTA partial layout (floats), identical to the FFMA kernel: dW2 [H*H] | dW3 [H*H] | db2 | db3 | dwc | dwr
constexpr int EDGE_PARTIAL = 2 * ENF_H * ENF_H + 4 * ENF_H;

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_edge_bwd_tc(const unsigned char* __restrict__ geom, const int* __restrict__ E_dev,
              const float* __restrict__ P, const float* __restrict__ S, const float* __restrict__ W1, int e1,
              const float* __restrict__ b2, const float* __restrict__ b3, const float* __restrict__ wc,
              const unsigned char* __restrict__ wimg, const float* __restrict__ dagg,
              float* __restrict__ runs, float* __restrict__ dz1, float* __restrict__ dd_out,
              float* __restrict__ partial) {
    using L = SmemB<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* Wimg = sm + L::w_off;
    unsigned char* XB = sm + L::x_off;         // x1 image [e][k], later x2^T image [k][e], later x1 again
    unsigned char* ZB = sm + L::z_off;         // dz3^T then dz2^T image [n][e]
    TileInfoB* tib = reinterpret_cast<TileInfoB*>(sm + L::t_off);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(sm + L::bar_off);
    uint64_t* bar_mma = bar_w + 1;
    uint64_t* bar_wg = bar_w + 2;          // the weight-gradient MMAs of a phase have completed (operands reusable)
    uint64_t* bar_g1 = bar_w + 3;          // the first GEMM of a tile is issued one tile ahead: its own barrier
    uint64_t* bar_geom = bar_w + 4;        // tile records arriving by TMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 5);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = w & 3, cg = w >> 2;
    const int n = 32 * q + lane;           // hidden unit == TMEM lane
    const int ec = 16 * cg;                // first tile edge of this thread's 16 columns

    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_wg, 1);
        tc::mbar_init(bar_g1, 1);
        tc::mbar_init(bar_geom, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    const float b2n = b2[n], b3n = b3[n], wcn = wc[n], wrn = W1[n * e1 + e1 - 1];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        tc::mbar_expect_tx(bar_w, L::NW * tc::IMG_BYTES);
        for (int i = 0; i < L::NW; ++i) {
            const int src = SPLIT ? i : 2 * i;      // global order: W2_hi, W2_lo, W3_hi, W3_lo
            tc::bulk_g2s(Wimg + (size_t)i * tc::IMG_BYTES, wimg + (size_t)src * tc::IMG_BYTES, tc::IMG_BYTES, bar_w);
        }
    }
    tc::mbar_wait(bar_w, 0);

    const uint32_t xb = tc::smem_u32(XB), zb = tc::smem_u32(ZB);
    const uint32_t w2 = tc::smem_u32(Wimg), w3 = tc::smem_u32(Wimg + (size_t)(SPLIT ? 2 : 1) * tc::IMG_BYTES);
    const uint32_t WLO = tc::IMG_BYTES, ALO = ACT_IMG;
    const uint32_t id_kmn64 = tc::make_idesc(false, true, 64);
    const uint32_t id_mm64 = tc::make_idesc(true, true, 64);
    const uint32_t id_kk128 = tc::make_idesc(false, false, 128);
    // base descriptors (K-major: LBO unused, SBO = 1024; MN-major: LBO = distance between 64-wide M/N blocks)
    const uint64_t dW2k = tc::make_desc(w2, 16, 1024), dW3k = tc::make_desc(w3, 16, 1024);
    const uint64_t dW2m = tc::make_desc(w2, tc::BLK_BYTES, 1024), dW3m = tc::make_desc(w3, tc::BLK_BYTES, 1024);
    const uint64_t dXk = tc::make_desc(xb, 16, 1024);                  // x1^T / x2^T [k][e] read K-major (K = e)
    const uint64_t dXTm = tc::make_desc(xb, tc::BLK_BYTES, 1024);      // x1^T / x2^T [k][e] read with rows = K = k
    const uint64_t dZk = tc::make_desc(zb, 16, 1024), dZm = tc::make_desc(zb, tc::BLK_BYTES, 1024);
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    uint32_t parity = 0;
    float gb2 = 0.f, gb3 = 0.f, gwc = 0.f, gwr = 0.f;
    bool first_tile = true;
    const int E = E_dev[0];
    const int tiles = (E + TE - 1) / TE;

    // z1 = P[row] + S[col] + w_r r for this thread's (hidden unit, 16 edges); row reads coalesce over the 32 hidden
    // units of a warp.  The values stay in registers for the whole tile (x1^T is needed twice) and are gathered
    // one tile ahead, behind the last MMA of the previous tile.
    auto load_z1 = [&](const TileInfoB& ti, float (&z)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int m = ec + j;
            z[j] = fmaf(wrn, ti.r[m], __ldg(P + (int64_t)ti.row[m] * ENF_H + n) + __ldg(S + (int64_t)ti.col[m] * ENF_H + n));
        }
    };
    // x1^T = silu(z1)^T into XB as the [hidden][edge] image; ds1 != nullptr also returns silu'(z1)
    auto put_x1 = [&](const float (&z1)[16], float* ds1) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = z1[8 * ch + j];
                const float sg = tc::sigmoid_sfu(z);
                x[j] = z * sg;
                if (ds1) ds1[8 * ch + j] = fmaf(x[j], 1.0f - sg, sg);       // silu'(z) = s + z s (1 - s)
            }
            store8<SPLIT>(XB, t_off_(n, 2 * cg + ch), x);
        }
    };
    // the record of one tile (k_edge_geom_bwd) into a TileInfoB, asynchronously, by one thread
    auto fetch_tile = [&](TileInfoB& ti, int tile) {
        tc::mbar_expect_tx(bar_geom, GEOM_BYTES);
        tc::bulk_g2s(&ti, geom + (int64_t)tile * GEOM_BYTES, GEOM_BYTES, bar_geom);
    };
    uint32_t parity_geom = 0;
    auto wait_tile = [&]() {
        tc::mbar_wait(bar_geom, parity_geom);
        parity_geom ^= 1;
    };
    auto issue_mma = [&](auto&& body, bool commit = true) {          // one thread issues
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            body();
            if (commit) tc::mma_commit(bar_mma);
        }
    };
    auto wait_mma = [&]() {                      // everybody waits for completion
        tc::mbar_wait(bar_mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
    };
    uint32_t parity_wg = 0;
    auto wait_wgrad = [&]() {                    // dgrad first, wgrad behind it: only operand reuse waits for the wgrad
        tc::mbar_wait(bar_wg, parity_wg);
        parity_wg ^= 1;
    };

    // Padding edges of the last tile are treated as self-edges of atom 0 with ds = 0 and dagg masked to 0: their
    // activations are finite and every gradient quantity that touches them is exactly zero, so the epilogues carry
    // no per-element validity selects (only stores and the dagg gather are predicated).
    // Software pipeline over the CTA's tiles (t, t+1, t+2 = this CTA's consecutive tiles):
    //   geometry(t+2) is computed behind the second MMA of tile t, the z1 gather of t+1 behind the last one;
    //   x1^T(t+1) is written and T1(t+1) = W2 x1^T issued (into the other T1 accumulator) before the last epilogue
    //   of tile t, so that MMA runs under the epilogue.
    uint32_t t1col = T1_COL, parity_g1 = 0;
    auto issue_g1 = [&](uint32_t col) {
        issue_mma([&]() {
            tc::issue_gemm_t<SPLIT, 8, tc::OffK128, tc::OffMN>(tmem + col, dW2k, WLO, dXTm, ALO, id_kmn64, false);
            tc::mma_commit(bar_g1);
        }, false);
    };
    int cur = 0;
    float z1[16];
    const int stride = gridDim.x;
    if ((int)blockIdx.x < tiles) {
        if (tid == 0) fetch_tile(tib[0], blockIdx.x);
        wait_tile();
        if ((int)blockIdx.x + stride < tiles) {
            if (tid == 0) fetch_tile(tib[1], blockIdx.x + stride);
            wait_tile();
        }
        load_z1(tib[0], z1);
        put_x1(z1, nullptr);
        issue_g1(t1col);
    }
    int tile_no = -1;
    for (int tile = blockIdx.x; tile < tiles; tile += stride) {
        ++tile_no;
        STAMP(0);
        const int e0 = tile * TE;
        TileInfoB& ti = tib[cur];
        TileInfoB& tn = tib[cur == 2 ? 0 : cur + 1];
        TileInfoB& tnn = tib[cur == 0 ? 2 : cur - 1];
        const int next = tile + stride, next2 = next + stride;
        // the next tile's record was requested one iteration ago (the first two in the prologue).  Everybody observes
        // its arrival HERE, before the barrier in front of the next request: an mbarrier must not run two phases
        // ahead of a waiter.
        if (tile != (int)blockIdx.x && next < tiles) wait_tile();
        // ---- T1 = W2 x1^T (issued one tile ahead)
        tc::mbar_wait(bar_g1, parity_g1);
        parity_g1 ^= 1;
        tc::fence_after_sync();
        STAMP(1);
        float dsl2[16];                 // silu'(z2), consumed two phases later
        {
            tc::tmem_ld16(lane_base + t1col + ec, dsl2);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int jj = 8 * ch + j;
                    const float z = dsl2[jj] + b2n;
                    const float sg = tc::sigmoid_sfu(z);
                    x[j] = z * sg;
                    dsl2[jj] = fmaf(x[j], 1.0f - sg, sg);
                }
                store8<SPLIT>(XB, t_off_(n, 2 * cg + ch), x);          // x2^T image [k][e] over x1^T
            }
        }
        STAMP(2);
        // ---- T2 = W3 x2^T
        issue_mma([&]() { tc::issue_gemm_t<SPLIT, 8, tc::OffK128, tc::OffMN>(tmem + T2_COL, dW3k, WLO, dXTm, ALO, id_kmn64, false); });
        STAMP(3);
        if (next2 < tiles && tid == 0) fetch_tile(tnn, next2);        // two tiles ahead; waited for before its first use
        wait_mma();
        STAMP(4);
        {
            float v[16];
            tc::tmem_ld16(lane_base + T2_COL + ec, v);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int jj = 8 * ch + j;
                    const float z = v[jj] + b3n;
                    const float sg = tc::sigmoid_sfu(z);
                    const float ds = ti.ds[ec + jj];
                    const float x3 = z * sg;
                    gwc = fmaf(ds, x3, gwc);
                    const float dz = ds * wcn * fmaf(x3, 1.0f - sg, sg);
                    gb3 += dz;
                    x[j] = dz;
                }
                store8<SPLIT>(ZB, t_off_(n, 2 * cg + ch), x);          // dz3^T image [n][e]
            }
        }
        STAMP(5);
        // ---- TW3 += dz3^T x2 ; T2 = W3^T dz3^T   (while the tensor pipe runs: gather dagg)
        issue_mma([&]() {
            tc::issue_gemm_t<SPLIT, 8, tc::OffMN, tc::OffMN>(tmem + T2_COL, dW3m, WLO, dZm, ALO, id_mm64, false);
            tc::mma_commit(bar_mma);             // the epilogue only needs the dgrad ...
            tc::issue_gemm_t<SPLIT, 4, tc::OffK, tc::OffK>(tmem + TW3_COL, dZk, ALO, dXk, ALO, id_kk128, !first_tile);
            tc::mma_commit(bar_wg);              // ... the wgrad finishes behind it
        }, false);
        STAMP(6);
        float da[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            da[j] = ti.valid[ec + j] ? __ldg(dagg + (int64_t)ti.row[ec + j] * ENF_H + n) : 0.f;
        STAMP(7);
        wait_mma();
        STAMP(8);
        {
            float v[16];
            tc::tmem_ld16(lane_base + T2_COL + ec, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = (v[j] + da[j]) * dsl2[j];
                gb2 += v[j];
            }
            wait_wgrad();                        // dz3^T / x2^T are still being read by the TW3 MMAs until here
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[8 * ch + j];
                store8<SPLIT>(ZB, t_off_(n, 2 * cg + ch), x);          // dz2^T image [n][e]
            }
        }
        STAMP(9);
        float ds1[16];
        put_x1(z1, ds1);                                          // x2^T is dead: rebuild x1^T, keep silu'(z1)
        STAMP(10);
        // z1's registers are free: request the next tile's S rows before the barrier in front of the MMA issue
        if (next < tiles) {
#pragma unroll
            for (int j = 0; j < 16; ++j) z1[j] = __ldg(S + (int64_t)tn.col[ec + j] * ENF_H + n);
        }
        // ---- TW2 += dz2^T x1 ; T1 = W2^T dz2^T   
        issue_mma([&]() {
            tc::issue_gemm_t<SPLIT, 8, tc::OffMN, tc::OffMN>(tmem + t1col, dW2m, WLO, dZm, ALO, id_mm64, false);
            tc::mma_commit(bar_mma);
            tc::issue_gemm_t<SPLIT, 4, tc::OffK, tc::OffK>(tmem + TW2_COL, dZk, ALO, dXk, ALO, id_kk128, !first_tile);
            tc::mma_commit(bar_wg);              // waited for before x1^T is rewritten (below, or after the last tile)
        }, false);
        first_tile = false;
        STAMP(11);
        if (next < tiles) {                                             // rest of the next tile's gather, behind the MMAs
#pragma unroll
            for (int j = 0; j < 16; ++j)
                z1[j] = fmaf(wrn, tn.r[ec + j], __ldg(P + (int64_t)tn.row[ec + j] * ENF_H + n) + z1[j]);
        }
        STAMP(12);                              // next tile's gather, behind the MMAs
        wait_mma();
        STAMP(13);
        if (next < tiles) {                  // next tile's x1^T and its first MMA, which then runs under the epilogue
            wait_wgrad();
            put_x1(z1, nullptr);
            issue_g1(T1B_COL - t1col);
        }
        STAMP(14);
        {
            float v[16];
            tc::tmem_ld16(lane_base + t1col + ec, v);
            // dz1 = dx1 * silu'(z1): stored per edge (for the column-grouped sum dS) and reduced over each row's
            // edges into per-run partials (dP, see segment.cu) by a thread-local running sum
            int rid = ((e0 + ec) >> 4) + ti.mis[ec];
            float acc = 0.f;
            const bool any = ti.valid[ec];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int m = ec + j;
                const float dz = v[j] * ds1[j];
                if (ti.valid[m]) dz1[(int64_t)(e0 + m) * ENF_H + n] = dz;
                if (j > 0 && ti.start[m]) {
                    runs[(int64_t)rid * ENF_H + n] = acc;
                    ++rid;
                    acc = 0.f;
                }
                acc += dz;
                gwr = fmaf(dz, ti.r[m], gwr);
                v[j] = wrn * dz;
            }
            if (any) runs[(int64_t)rid * ENF_H + n] = acc;
            const float t = warp_transpose_sum16(v, lane);
            if (lane < 16) ti.dr_part[q][ec + lane] = t;
        }
        __syncthreads();
        if (tid < TE && ti.valid[tid]) {
            const float dr2 = 2.0f * ((ti.dr_part[0][tid] + ti.dr_part[1][tid]) + (ti.dr_part[2][tid] + ti.dr_part[3][tid]));
#pragma unroll
            for (int c = 0; c < 3; ++c) dd_out[(int64_t)(e0 + tid) * 3 + c] = fmaf(dr2, ti.d[tid][c], ti.ddir[tid][c]);
        }
        STAMP(15);
        cur = cur == 2 ? 0 : cur + 1;
        t1col = T1B_COL - t1col;
    }
    // ---- per-CTA partials: weight gradients from TMEM, vector gradients combined over the 4 edge groups
    if (!first_tile) wait_wgrad();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    float* my = partial + (int64_t)blockIdx.x * EDGE_PARTIAL;
    {
        float v[32];
#pragma unroll 1
        for (int mat = 0; mat < 2; ++mat) {
            if (first_tile) {
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
    float* red = reinterpret_cast<float*>(XB);          // [4 kinds][4 cg][128]
    red[(0 * 4 + cg) * ENF_H + n] = gb2;
    red[(1 * 4 + cg) * ENF_H + n] = gb3;
    red[(2 * 4 + cg) * ENF_H + n] = gwc;
    red[(3 * 4 + cg) * ENF_H + n] = gwr;
    __syncthreads();
    if (tid < 4 * ENF_H) {
        const int a = tid / ENF_H, k = tid % ENF_H;
        my[2 * ENF_H * ENF_H + tid] = (red[(a * 4 + 0) * ENF_H + k] + red[(a * 4 + 1) * ENF_H + k]) +
                                      (red[(a * 4 + 2) * ENF_H + k] + red[(a * 4 + 3) * ENF_H + k]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

int enf_edge_reduce_partials(const float* partial, int n_cta, float* lgrad, int nf, cudaStream_t st);

// bytes of the per-tile records k_edge_geom_bwd writes for a capacity of E_cap edges (16-byte aligned buffer)
int64_t enf_edge_bwd_geom_bytes(int E_cap) { return ((int64_t)(E_cap + TE - 1) / TE + 1) * GEOM_BYTES; }

int enf_edge_bwd_tc(int mode, const int* row, const int* col, const int* rowptr, const int* E_dev, int E_cap,
                    const float* pos, const float* box, const float* P, const float* S, const float* lp,
                    const unsigned char* wimg, int nf, const float* s_saved, const float* dagg, const float* dF,
                    float coords_weight, const int* mis, float* runs, float* dz1, float* dd, float* lgrad,
                    float* partial, unsigned char* geom, cudaStream_t st) {
    if (E_cap == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const int grid = enf_num_sms();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_edge_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemB<true>::total);
        cudaFuncSetAttribute(k_edge_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemB<false>::total);
        attr = true;
    }
    const int slots = (E_cap + TE - 1) / TE * TE;
    int ggrid = (slots + 255) / 256;
    if (ggrid > 8 * enf_num_sms()) ggrid = 8 * enf_num_sms();
    enf_count_launch(), k_edge_geom_bwd<<<ggrid, 256, 0, st>>>(row, col, rowptr, E_dev, pos, box, s_saved, dF, coords_weight, mis, geom);
    if (mode == 1)
        enf_count_launch(), k_edge_bwd_tc<true><<<grid, THREADS, SmemB<true>::total, st>>>(
            geom, E_dev, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], wimg, dagg,
            runs, dz1, dd, partial);
    else
        enf_count_launch(), k_edge_bwd_tc<false><<<grid, THREADS, SmemB<false>::total, st>>>(
            geom, E_dev, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], wimg, dagg,
            runs, dz1, dd, partial);
    ENF_CHECK_LAUNCH();
    return enf_edge_reduce_partials(partial, grid, lgrad, nf, st);
}
