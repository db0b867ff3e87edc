// node_model of one EGCL (enflow/nn/egcl.py:27-30,65-69) on tcgen05, forward and backward, 128 nodes per tile.
//
// Accumulators are kept TRANSPOSED like in the edge kernels (TMEM lane = hidden unit, column = node), so the
// per-node rows of z4 / agg / dagg are read and written as 128-byte warp transactions and the operand images
// are written as contiguous 8-node chunks.  The nf-wide heads (G = W5 silu(z4), dh = W4[:, :nf]^T dz4) are
// N = 16 MMAs whose accumulator has the node on the TMEM lane.
//
//   forward : Tz [k][i]  = W4 [agg; h]^T          A = W4 image [k][j] (K-major)        B = IN image [j][i] (MN-major)
//             Tg [i][c]  = x4 W5^T                A = x4^T image [k][i] (MN-major)     B = W5 image [c][k] (K-major)
//   backward: Ta [jj][i] = W4A^T dz4^T            A = W4A image [k][jj] (MN-major)     B = dz4^T image [k][i] (MN-major)
//             Th [i][c]  = dz4 W4H                A = dz4^T image [k][i] (MN-major)    B = W4H image [c][k] (K-major)
//             TW [k][j] += dz4^T [agg; h]         A = dz4^T image [k][i] (K-major)     B = IN image [j][i] (K-major)
// IN = [agg (128 rows); h (nf rows, padded to 16)] x 128 nodes; the W4 images use the same input order.
// TW stays in TMEM for the life of the CTA; per-CTA partials are combined in CTA order (deterministic).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int THREADS = 512;
constexpr int TN = 128;                          // nodes per tile
constexpr int IN_ROWS = 144;
constexpr int IN_BLK = IN_ROWS * 128;            // bytes of one 64-node block of the IN image
constexpr int IN_IMG = 2 * IN_BLK;
constexpr int S16_BLK = 16 * 128;                // 16-row images (W4H, W5): one 64-column block
constexpr int S16_IMG = 2 * S16_BLK;
constexpr int W4F_IMG = 3 * tc::BLK_BYTES;       // [128][192]: 144 input columns used

struct OffK144 { static constexpr uint32_t off(int ks) { return (uint32_t)((ks >> 2) * IN_BLK + (ks & 3) * 32); } };
struct OffK16 { static constexpr uint32_t off(int ks) { return (uint32_t)((ks >> 2) * S16_BLK + (ks & 3) * 32); } };

// byte offset of the 16-byte chunk holding columns [8 c16, 8 c16 + 8) of `row` in an image with `blk` bytes per block
__host__ __device__ inline uint32_t chunk_off(int row, int c16, int blk) {
    return (uint32_t)((c16 >> 3) * blk + row * 128 + (((c16 & 7) ^ (row & 7)) << 4));
}

template <bool SPLIT>
__device__ __forceinline__ void put8(unsigned char* img, uint32_t off, uint32_t lo_bytes, const float (&x)[8]) {
    if (SPLIT) {
        uint4 hi, lo;
        tc::split2(x[0], x[1], hi.x, lo.x);
        tc::split2(x[2], x[3], hi.y, lo.y);
        tc::split2(x[4], x[5], hi.z, lo.z);
        tc::split2(x[6], x[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(img + off) = hi;
        *reinterpret_cast<uint4*>(img + lo_bytes + off) = lo;
    } else {
        uint4 hi;
        hi.x = tc::pack_bf16(x[0], x[1]); hi.y = tc::pack_bf16(x[2], x[3]);
        hi.z = tc::pack_bf16(x[4], x[5]); hi.w = tc::pack_bf16(x[6], x[7]);
        *reinterpret_cast<uint4*>(img + off) = hi;
    }
}

// ---- weight images (global, per layer, behind the four edge images) ------------------------------------------
__global__ void __launch_bounds__(256) k_pack_node_tc(const float* __restrict__ W4, const float* __restrict__ W5, int nf,
                                                       unsigned char* __restrict__ img, int64_t param_stride,
                                                       int64_t img_stride) {
    const int D = nf + ENF_H;
    W4 += blockIdx.y * param_stride; W5 += blockIdx.y * param_stride; img += blockIdx.y * img_stride;     // blockIdx.y = layer
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float x[8];
    unsigned char* dst;
    uint32_t off, lo;
    if (idx < 128 * 16) {                          // W4A [k][jj]
        const int k = idx >> 4, ch = idx & 15;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = W4[(int64_t)k * D + nf + 8 * ch + i];
        dst = img + tc::NODE_W4A; off = tc::img_chunk_offset(k, ch); lo = tc::IMG_BYTES;
    } else if ((idx -= 128 * 16) < 16 * 16) {      // W4H [c][k]
        const int c = idx >> 4, ch = idx & 15;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = c < nf ? W4[(int64_t)(8 * ch + i) * D + c] : 0.f;
        dst = img + tc::NODE_W4H; off = chunk_off(c, ch, S16_BLK); lo = S16_IMG;
    } else if ((idx -= 16 * 16) < 128 * 24) {      // W4F [k][agg 0..127 | h 128..]
        const int k = idx / 24, ch = idx % 24;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = 8 * ch + i;
            x[i] = j < ENF_H ? W4[(int64_t)k * D + nf + j] : (j - ENF_H < nf ? W4[(int64_t)k * D + (j - ENF_H)] : 0.f);
        }
        dst = img + tc::NODE_W4F; off = chunk_off(k, ch, tc::BLK_BYTES); lo = W4F_IMG;
    } else if ((idx -= 128 * 24) < 16 * 16) {      // W5 [c][k]
        const int c = idx >> 4, ch = idx & 15;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = c < nf ? W5[c * ENF_H + 8 * ch + i] : 0.f;
        dst = img + tc::NODE_W5; off = chunk_off(c, ch, S16_BLK); lo = S16_IMG;
    } else {
        return;
    }
    put8<true>(dst, off, lo, x);
}
constexpr int PACK_ITEMS = 128 * 16 + 16 * 16 + 128 * 24 + 16 * 16;

// IN image rows 128.. (the h columns): 16 rows x 16 node chunks, one item per thread
template <bool SPLIT>
__device__ __forceinline__ void build_in_h(unsigned char* IN, const float* __restrict__ h, int t0, int N, int nf, int tid) {
    if (tid < 256) {
        const int c = tid >> 4, ch = tid & 15;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int node = t0 + 8 * ch + i;
            x[i] = (c < nf && node < N) ? __ldg(h + (int64_t)node * nf + c) : 0.f;
        }
        put8<SPLIT>(IN, chunk_off(ENF_H + c, ch, IN_BLK), IN_IMG, x);
    }
}

struct Bars {
    uint64_t w, mma, wg;
    uint32_t tmem_slot, pad;
};

// ================================================ forward =====================================================
template <bool SPLIT>
struct SmemF {
    static constexpr int NI = SPLIT ? 2 : 1;
    static constexpr size_t w4_off = 0;
    static constexpr size_t w5_off = (size_t)NI * W4F_IMG;
    static constexpr size_t in_off = w5_off + (size_t)NI * S16_IMG;       // also the x4^T image after the first GEMM
    static constexpr size_t bar_off = in_off + (size_t)NI * IN_IMG;
    static constexpr size_t total = bar_off + sizeof(Bars) + 1024;
};
constexpr uint32_t F_TZ = 0, F_TG = 128, F_COLS = 256;

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_node_post_fwd_tc(const float* __restrict__ h, const float* __restrict__ agg, int N, int nf,
                   const float* __restrict__ b4, const float* __restrict__ b5, const unsigned char* __restrict__ wimg,
                   float* __restrict__ z4, float* __restrict__ G) {
    using L = SmemF<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* W4I = sm + L::w4_off;
    unsigned char* W5I = sm + L::w5_off;
    unsigned char* IN = sm + L::in_off;
    Bars& bars = *reinterpret_cast<Bars*>(sm + L::bar_off);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = w & 3, cg = w >> 2;
    const int n = 32 * q + lane;
    if (tid == 0) {
        tc::mbar_init(&bars.w, 1);
        tc::mbar_init(&bars.mma, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(&bars.tmem_slot, F_COLS);
    const float b4n = b4[n];
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;
    if (tid == 0) {
        tc::mbar_expect_tx(&bars.w, L::NI * (W4F_IMG + S16_IMG));
        for (int i = 0; i < L::NI; ++i) {
            tc::bulk_g2s(W4I + (size_t)i * W4F_IMG, wimg + tc::NODE_W4F + (size_t)i * W4F_IMG, W4F_IMG, &bars.w);
            tc::bulk_g2s(W5I + (size_t)i * S16_IMG, wimg + tc::NODE_W5 + (size_t)i * S16_IMG, S16_IMG, &bars.w);
        }
    }
    const uint32_t id_z = tc::make_idesc(false, true, 128);
    const uint32_t id_g = tc::make_idesc(true, false, 16);
    const uint64_t dW4k = tc::make_desc(tc::smem_u32(W4I), 16, 1024);
    const uint64_t dW5k = tc::make_desc(tc::smem_u32(W5I), 16, 1024);
    const uint64_t dINm = tc::make_desc(tc::smem_u32(IN), IN_BLK, 1024);
    const uint64_t dX4m = tc::make_desc(tc::smem_u32(IN), tc::BLK_BYTES, 1024);
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    uint32_t parity = 0;
    auto run_mma = [&](auto&& body) {
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            body();
            tc::mma_commit(&bars.mma);
        }
        tc::mbar_wait(&bars.mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
    };
    bool weights_ready = false;
    const int tiles = (N + TN - 1) / TN;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t0 = tile * TN;
        build_in_h<SPLIT>(IN, h, t0, N, nf, tid);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int node = t0 + 32 * cg + 8 * ch + j;
                x[j] = node < N ? __ldg(agg + (int64_t)node * ENF_H + n) : 0.f;
            }
            put8<SPLIT>(IN, chunk_off(n, 4 * cg + ch, IN_BLK), IN_IMG, x);
        }
        if (!weights_ready) {
            tc::mbar_wait(&bars.w, 0);
            weights_ready = true;
        }
        run_mma([&]() { tc::issue_gemm_t<SPLIT, 9, tc::OffK128, tc::OffMN>(tmem + F_TZ, dW4k, W4F_IMG, dINm, IN_IMG, id_z, false); });
        {
            float v[32];
            tc::tmem_ld32(lane_base + F_TZ + 32 * cg, v);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int node = t0 + 32 * cg + 8 * ch + j;
                    const float z = v[8 * ch + j] + b4n;
                    if (z4 != nullptr && node < N) z4[(int64_t)node * ENF_H + n] = z;      // kept for the backward pass only
                    x[j] = z * tc::sigmoid_sfu(z);
                }
                put8<SPLIT>(IN, tc::img_chunk_offset(n, 4 * cg + ch), tc::IMG_BYTES, x);        // x4^T over the IN image
            }
        }
        run_mma([&]() { tc::issue_gemm_t<SPLIT, 8, tc::OffMN, OffK16>(tmem + F_TG, dX4m, tc::IMG_BYTES, dW5k, S16_IMG, id_g, false); });
        if (cg == 0) {
            float u[16];
            tc::tmem_ld16(lane_base + F_TG, u);
            const int node = t0 + n;
            if (node < N) {
#pragma unroll
                for (int c = 0; c < ENF_MAX_NF; ++c)
                    if (c < nf) G[(int64_t)node * nf + c] = u[c] + __ldg(b5 + c);
            }
        }
        tc::fence_before_sync();
        __syncthreads();          // the next tile rewrites the image and the accumulators
        tc::fence_after_sync();
    }
    if (!weights_ready) tc::mbar_wait(&bars.w, 0);      // never leave with a bulk copy in flight
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, F_COLS);
}

// ================================================ backward ====================================================
template <bool SPLIT>
struct SmemB {
    static constexpr int NI = SPLIT ? 2 : 1;
    static constexpr size_t w4a_off = 0;
    static constexpr size_t w4h_off = (size_t)NI * tc::IMG_BYTES;
    static constexpr size_t in_off = w4h_off + (size_t)NI * S16_IMG;
    static constexpr size_t z_off = in_off + (size_t)NI * IN_IMG;
    static constexpr size_t dg_off = z_off + (size_t)NI * tc::IMG_BYTES;
    static constexpr size_t bar_off = dg_off + sizeof(float) * TN * ENF_MAX_NF;
    static constexpr size_t total = bar_off + sizeof(Bars) + 1024;
};
constexpr uint32_t B_TA = 0, B_TH = 128, B_TW = 256, B_COLS = 512;

// per-CTA partial (floats), the layout of node_post.cu: dW4 [H*D] native [k][j] | db4 [H] | dW5 [nf*H] | db5 [nf]
template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_node_post_bwd_tc(const float* __restrict__ h, const float* __restrict__ agg, const float* __restrict__ z4,
                   const float* __restrict__ dG, int N, int nf, const float* __restrict__ W5,
                   const unsigned char* __restrict__ wimg, float* __restrict__ dagg, float* __restrict__ dh,
                   float* __restrict__ partial) {
    using L = SmemB<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* WA = sm + L::w4a_off;
    unsigned char* WH = sm + L::w4h_off;
    unsigned char* IN = sm + L::in_off;
    unsigned char* ZB = sm + L::z_off;
    float* dg_s = reinterpret_cast<float*>(sm + L::dg_off);          // [TN][8]
    Bars& bars = *reinterpret_cast<Bars*>(sm + L::bar_off);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = w & 3, cg = w >> 2;
    const int n = 32 * q + lane;
    const int D = nf + ENF_H;
    if (tid == 0) {
        tc::mbar_init(&bars.w, 1);
        tc::mbar_init(&bars.mma, 1);
        tc::mbar_init(&bars.wg, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(&bars.tmem_slot, B_COLS);
    float w5[ENF_MAX_NF], gw5[ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        w5[c] = c < nf ? W5[c * ENF_H + n] : 0.f;
        gw5[c] = 0.f;
    }
    float gb4 = 0.f, gb5 = 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = bars.tmem_slot;
    if (tid == 0) {
        tc::mbar_expect_tx(&bars.w, L::NI * (tc::IMG_BYTES + S16_IMG));
        for (int i = 0; i < L::NI; ++i) {
            tc::bulk_g2s(WA + (size_t)i * tc::IMG_BYTES, wimg + tc::NODE_W4A + (size_t)i * tc::IMG_BYTES, tc::IMG_BYTES, &bars.w);
            tc::bulk_g2s(WH + (size_t)i * S16_IMG, wimg + tc::NODE_W4H + (size_t)i * S16_IMG, S16_IMG, &bars.w);
        }
    }
    const uint32_t id_a = tc::make_idesc(true, true, 128);
    const uint32_t id_h = tc::make_idesc(true, false, 16);
    const uint32_t id_w = tc::make_idesc(false, false, IN_ROWS);
    const uint64_t dWAm = tc::make_desc(tc::smem_u32(WA), tc::BLK_BYTES, 1024);
    const uint64_t dWHk = tc::make_desc(tc::smem_u32(WH), 16, 1024);
    const uint64_t dZm = tc::make_desc(tc::smem_u32(ZB), tc::BLK_BYTES, 1024);
    const uint64_t dZk = tc::make_desc(tc::smem_u32(ZB), 16, 1024);
    const uint64_t dINk = tc::make_desc(tc::smem_u32(IN), 16, 1024);
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    uint32_t parity = 0, parity_wg = 0;
    bool first_tile = true, weights_ready = false;

    const int tiles = (N + TN - 1) / TN;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t0 = tile * TN;
        for (int idx = tid; idx < TN * ENF_MAX_NF; idx += THREADS) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            dg_s[idx] = (t0 + t < N && c < nf) ? dG[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        build_in_h<SPLIT>(IN, h, t0, N, nf, tid);
        __syncthreads();
        // dz4 = (W5^T dG) silu'(z4); rows of padding nodes have dG = 0, hence dz4 = 0
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float xz[8], xa[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = 32 * cg + 8 * ch + j;
                const int node = t0 + t;
                float z = 0.f, a = 0.f;
                if (node < N) {
                    z = __ldg(z4 + (int64_t)node * ENF_H + n);
                    a = __ldg(agg + (int64_t)node * ENF_H + n);
                }
                const float sg = tc::sigmoid_sfu(z);
                const float x = z * sg;
                const float4 g0 = *reinterpret_cast<const float4*>(dg_s + t * ENF_MAX_NF);
                const float4 g1 = *reinterpret_cast<const float4*>(dg_s + t * ENF_MAX_NF + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                float dx = 0.f;
#pragma unroll
                for (int c = 0; c < ENF_MAX_NF; ++c) {
                    dx = fmaf(w5[c], g[c], dx);
                    gw5[c] = fmaf(g[c], x, gw5[c]);
                }
                const float dz = dx * fmaf(x, 1.0f - sg, sg);
                gb4 += dz;
                xz[j] = dz;
                xa[j] = a;
            }
            put8<SPLIT>(ZB, tc::img_chunk_offset(n, 4 * cg + ch), tc::IMG_BYTES, xz);
            put8<SPLIT>(IN, chunk_off(n, 4 * cg + ch, IN_BLK), IN_IMG, xa);
        }
        if (n < ENF_MAX_NF) {
#pragma unroll 8
            for (int j = 0; j < 32; ++j) gb5 += dg_s[(32 * cg + j) * ENF_MAX_NF + n];
        }
        if (!weights_ready) {
            tc::mbar_wait(&bars.w, 0);
            weights_ready = true;
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            tc::issue_gemm_t<SPLIT, 8, tc::OffMN, tc::OffMN>(tmem + B_TA, dWAm, tc::IMG_BYTES, dZm, tc::IMG_BYTES, id_a, false);
            tc::issue_gemm_t<SPLIT, 8, tc::OffMN, OffK16>(tmem + B_TH, dZm, tc::IMG_BYTES, dWHk, S16_IMG, id_h, false);
            tc::mma_commit(&bars.mma);
            tc::issue_gemm_t<SPLIT, 8, tc::OffK128, OffK144>(tmem + B_TW, dZk, tc::IMG_BYTES, dINk, IN_IMG, id_w, !first_tile);
            tc::mma_commit(&bars.wg);
        }
        first_tile = false;
        tc::mbar_wait(&bars.mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
        {
            float v[32];
            tc::tmem_ld32(lane_base + B_TA + 32 * cg, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int node = t0 + 32 * cg + j;
                if (node < N) dagg[(int64_t)node * ENF_H + n] = v[j];
            }
        }
        if (cg == 0) {
            float u[16];
            tc::tmem_ld16(lane_base + B_TH, u);
            const int node = t0 + n;
            if (node < N) {
#pragma unroll
                for (int c = 0; c < ENF_MAX_NF; ++c)
                    if (c < nf) dh[(int64_t)node * nf + c] += u[c];
            }
        }
        tc::mbar_wait(&bars.wg, parity_wg);          // the images are rewritten by the next tile
        parity_wg ^= 1;
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    if (!weights_ready) tc::mbar_wait(&bars.w, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    // ---- per-CTA partials
    float* p = partial + (int64_t)blockIdx.x * ((int64_t)ENF_H * D + ENF_H + nf * ENF_H + nf);
    {
        float v[32];
        if (first_tile) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
            tc::tmem_ld32(lane_base + B_TW + 32 * cg, v);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) p[(int64_t)n * D + nf + 32 * cg + j] = v[j];
        if (cg == 0) {
            float u[16];
            if (first_tile) {
#pragma unroll
                for (int j = 0; j < 16; ++j) u[j] = 0.f;
            } else {
                tc::tmem_ld16(lane_base + B_TW + ENF_H, u);
            }
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c)
                if (c < nf) p[(int64_t)n * D + c] = u[c];
        }
    }
    p += (int64_t)ENF_H * D;
    float* red = reinterpret_cast<float*>(ZB);          // [1 + nf rows][4 node groups][128], then db5 [4][8]
    red[(0 * 4 + cg) * ENF_H + n] = gb4;
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c)
        if (c < nf) red[((1 + c) * 4 + cg) * ENF_H + n] = gw5[c];
    float* red5 = red + (1 + ENF_MAX_NF) * 4 * ENF_H;
    if (n < ENF_MAX_NF) red5[cg * ENF_MAX_NF + n] = gb5;
    __syncthreads();
    for (int idx = tid; idx < (1 + nf) * ENF_H; idx += THREADS) {
        const int r = idx / ENF_H, k = idx % ENF_H;
        p[idx] = (red[(r * 4 + 0) * ENF_H + k] + red[(r * 4 + 1) * ENF_H + k]) +
                 (red[(r * 4 + 2) * ENF_H + k] + red[(r * 4 + 3) * ENF_H + k]);
    }
    if (tid < nf)
        p[(1 + nf) * ENF_H + tid] = (red5[0 * ENF_MAX_NF + tid] + red5[1 * ENF_MAX_NF + tid]) +
                                    (red5[2 * ENF_MAX_NF + tid] + red5[3 * ENF_MAX_NF + tid]);
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, B_COLS);
}

int tc_grid(int N) {
    const int tiles = (N + TN - 1) / TN;
    const int cap = enf_num_sms();
    return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

}  // namespace

int enf_node_post_reduce(const float* partial, int n_cta, int nf, float* lgrad, cudaStream_t st);

int enf_node_tc_pack(const float* lp0, int nf, int L, int64_t param_stride, unsigned char* img0, int64_t img_stride,
                     cudaStream_t st) {
    const EgclOffsets o = enf_egcl_offsets(nf);
    enf_count_launch(), k_pack_node_tc<<<dim3((PACK_ITEMS + 255) / 256, L), 256, 0, st>>>(lp0 + o.off[P_W4], lp0 + o.off[P_W5], nf,
                                                                                      img0, param_stride, img_stride);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// The node kernels see N rows where the edge kernels see E: they always run the bf16x3 split (fp32-accurate), also
// when the edge MLP runs in plain bf16 (mode 2) -- there is nothing to gain from dropping precision here.
int enf_node_post_fwd_tc(int mode, const float* h, const float* agg, int N, int nf, const float* lp,
                         const unsigned char* wimg, float* z4, float* G, cudaStream_t st) {
    (void)mode;
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_node_post_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemF<true>::total);
        attr = true;
    }
    enf_count_launch(), k_node_post_fwd_tc<true><<<tc_grid(N), THREADS, SmemF<true>::total, st>>>(
        h, agg, N, nf, lp + o.off[P_B4], lp + o.off[P_B5], wimg, z4, G);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_node_post_bwd_tc(int mode, const float* h, const float* agg, const float* z4, const float* dG, int N, int nf,
                         const float* lp, const unsigned char* wimg, float* dagg, float* dh, float* lgrad,
                         float* partial, cudaStream_t st, cudaStream_t st_red) {
    (void)mode;
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_node_post_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemB<true>::total);
        attr = true;
    }
    const int grid = tc_grid(N);
    enf_count_launch(), k_node_post_bwd_tc<true><<<grid, THREADS, SmemB<true>::total, st>>>(
        h, agg, z4, dG, N, nf, lp + o.off[P_W5], wimg, dagg, dh, partial);
    ENF_CHECK_LAUNCH();
    enf_chain(st, st_red);                  // the partial reduction runs beside whatever follows on st
    return enf_node_post_reduce(partial, grid, nf, lgrad, st_red);
}
