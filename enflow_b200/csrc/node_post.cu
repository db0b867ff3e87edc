// node_model of one EGCL (enflow/nn/egcl.py:27-30,65-69): G = W5 silu(W4 [h; agg] + b4) + b5, forward and backward.
// Register-tiled FFMA GEMMs over tiles of 32 nodes (these layers see N rows, not E: ~2 % of the FLOPs).
//   forward : thread = 4 nodes x 4 hidden units, inner loop over the 128+nf inputs, W4^T rows stream through L1
//   backward: dz4 in shared memory; dagg = dz4 W4[:, nf:] (same tiling); dW4 += dz4^T [h; agg] accumulated in
//             68 registers per thread across all tiles of the CTA (static tile schedule), then per-CTA partials
//             combined in CTA order: deterministic.
#include "common.cuh"

namespace {

constexpr int PT = 32;          // nodes per tile
constexpr int TPB = 256;
constexpr int DP = 136;         // padded input width (nf + 128 <= 136), float4-aligned rows
constexpr int JS = 9;           // backward: input columns per warp in the wgrad (16 warps x 9 = 144)
constexpr int BT = 512;         // backward: threads per CTA (16 warps hide the shared-memory latency of the wgrad loop)
constexpr int BPT = 64;         // backward: nodes per tile
constexpr int BDP = 144;        // backward: padded input width

template <int NODES, int STRIDE, int NTHREADS>
__device__ __forceinline__ void load_inputs(float* in_s, const float* __restrict__ h, const float* __restrict__ agg,
                                            int t0, int N, int nf) {
    for (int idx = threadIdx.x; idx < NODES * STRIDE; idx += NTHREADS) {
        const int t = idx / STRIDE, j = idx - t * STRIDE;
        float v = 0.f;
        if (t0 + t < N) {
            if (j < nf) v = h[(int64_t)(t0 + t) * nf + j];
            else if (j < nf + ENF_H) v = agg[(int64_t)(t0 + t) * ENF_H + (j - nf)];
        }
        in_s[idx] = v;
    }
}

__global__ void __launch_bounds__(TPB, 2) k_node_post_fwd(const float* __restrict__ h, const float* __restrict__ agg,
                                                           int N, int nf, const float* __restrict__ W4T,
                                                           const float* __restrict__ b4, const float* __restrict__ W5,
                                                           const float* __restrict__ b5, float* __restrict__ z4,
                                                           float* __restrict__ G) {
    extern __shared__ __align__(16) float smem_nf[];
    float* w_s = smem_nf;                         // W4^T [D][H], staged once per (persistent) CTA
    float* in_s = w_s + DP * ENF_H;               // [PT][DP]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int D = nf + ENF_H;
    for (int idx = threadIdx.x; idx < D * ENF_H / 4; idx += TPB)
        reinterpret_cast<float4*>(w_s)[idx] = __ldg(reinterpret_cast<const float4*>(W4T) + idx);
    const float4 bb = *reinterpret_cast<const float4*>(b4 + 4 * lane);
    const int tiles = (N + PT - 1) / PT;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t0 = tile * PT;
        __syncthreads();
        load_inputs<PT, DP, TPB>(in_s, h, agg, t0, N, nf);
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[t][0] = bb.x; acc[t][1] = bb.y; acc[t][2] = bb.z; acc[t][3] = bb.w; }
        const float* in0 = in_s + (4 * w) * DP;
#pragma unroll 4
        for (int j = 0; j < D; ++j) {
            const float4 wv = *reinterpret_cast<const float4*>(w_s + j * ENF_H + 4 * lane);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float a = in0[t * DP + j];
                acc[t][0] = fmaf(a, wv.x, acc[t][0]); acc[t][1] = fmaf(a, wv.y, acc[t][1]);
                acc[t][2] = fmaf(a, wv.z, acc[t][2]); acc[t][3] = fmaf(a, wv.w, acc[t][3]);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = t0 + 4 * w + t;
            if (i < N) *reinterpret_cast<float4*>(z4 + (int64_t)i * ENF_H + 4 * lane) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[t][c] = siluf_(acc[t][c]);
        }
        for (int c = 0; c < nf; ++c) {
            const float4 w5 = __ldg(reinterpret_cast<const float4*>(W5 + c * ENF_H) + lane);
            const float bc = b5[c];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float p = fmaf(w5.x, acc[t][0], fmaf(w5.y, acc[t][1], fmaf(w5.z, acc[t][2], w5.w * acc[t][3])));
                p = warp_sum(p);
                const int i = t0 + 4 * w + t;
                if (lane == 0 && i < N) G[(int64_t)i * nf + c] = p + bc;
            }
        }
    }
}

// per-CTA partial (floats): dW4 [H*D] (native [k][j]) | db4 [H] | dW5 [nf*H] | db5 [nf]
__global__ void __launch_bounds__(BT, 1) k_node_post_bwd(const float* __restrict__ h, const float* __restrict__ agg,
                                                           const float* __restrict__ z4, const float* __restrict__ dG,
                                                           int N, int nf, const float* __restrict__ W4,
                                                           const float* __restrict__ W4A, const float* __restrict__ W5,
                                                           float* __restrict__ dagg, float* __restrict__ dh,
                                                           float* __restrict__ partial) {
    extern __shared__ __align__(16) float smem_np[];
    float* in_s = smem_np;                       // [BPT][BDP]
    float* dz_s = in_s + BPT * BDP;              // [BPT][H]
    float* dg_s = dz_s + BPT * ENF_H;            // [BPT][MAX_NF]
    float* wa_s = dg_s + BPT * ENF_MAX_NF;       // W4[:, nf:] as [k][jj], staged once per CTA
    for (int idx = threadIdx.x; idx < ENF_H * ENF_H / 4; idx += BT)
        reinterpret_cast<float4*>(wa_s)[idx] = __ldg(reinterpret_cast<const float4*>(W4A) + idx);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int D = nf + ENF_H;
    float wacc[4][JS];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int t = 0; t < JS; ++t) wacc[c][t] = 0.f;
    float gb4[4] = {0.f, 0.f, 0.f, 0.f};
    float gw5[ENF_MAX_NF][4];
    float w5[ENF_MAX_NF][4];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        const float4 v = c < nf ? __ldg(reinterpret_cast<const float4*>(W5 + c * ENF_H) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        w5[c][0] = v.x; w5[c][1] = v.y; w5[c][2] = v.z; w5[c][3] = v.w;
        gw5[c][0] = gw5[c][1] = gw5[c][2] = gw5[c][3] = 0.f;
    }
    float gb5 = 0.f;      // lane c < nf: partial over this warp's nodes
    float w4h[ENF_MAX_NF][4];      // W4[4lane+c4][c] for the nf 'h' input columns (dh = W4[:, :nf]^T dz4)
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c)
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) w4h[c][c4] = c < nf ? __ldg(W4 + (int64_t)(4 * lane + c4) * D + c) : 0.f;
    const int tiles = (N + BPT - 1) / BPT;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int t0 = tile * BPT;
        __syncthreads();
        load_inputs<BPT, BDP, BT>(in_s, h, agg, t0, N, nf);
        for (int idx = threadIdx.x; idx < BPT * ENF_MAX_NF; idx += BT) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            dg_s[idx] = (t0 + t < N && c < nf) ? dG[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        __syncthreads();
        // ---- dz4 = (W5^T dG) * silu'(z4); thread = nodes 4w..4w+3 x hidden 4lane..4lane+3
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = t0 + 4 * w + t;
            float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < N) zv = __ldg(reinterpret_cast<const float4*>(z4 + (int64_t)i * ENF_H) + lane);
            const float z[4] = {zv.x, zv.y, zv.z, zv.w};
            float dz[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float sg = sigmoidf_(z[c4]);
                const float x = z[c4] * sg;
                float dx = 0.f;
#pragma unroll
                for (int c = 0; c < ENF_MAX_NF; ++c) {
                    const float g = dg_s[(4 * w + t) * ENF_MAX_NF + c];
                    dx = fmaf(w5[c][c4], g, dx);
                    gw5[c][c4] = fmaf(g, x, gw5[c][c4]);
                }
                dz[c4] = (i < N) ? dx * (sg * (1.0f + z[c4] * (1.0f - sg))) : 0.f;
                gb4[c4] += dz[c4];
            }
            *reinterpret_cast<float4*>(dz_s + (4 * w + t) * ENF_H + 4 * lane) = make_float4(dz[0], dz[1], dz[2], dz[3]);
            // dh[i][c] += sum_k W4[k][c] dz4[i][k]: lane-partial over its 4 hidden units, then a warp reduction
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) {
                if (c < nf) {
                    float pdh = fmaf(w4h[c][0], dz[0], fmaf(w4h[c][1], dz[1], fmaf(w4h[c][2], dz[2], w4h[c][3] * dz[3])));
                    pdh = warp_sum(pdh);
                    if (lane == 0 && i < N) dh[(int64_t)i * nf + c] += pdh;
                }
            }
            if (lane < nf) gb5 += dg_s[(4 * w + t) * ENF_MAX_NF + lane];
        }
        __syncthreads();
        // ---- dagg[i][jj] = sum_k W4A[k][jj] dz[i][k]; thread = 4 nodes x 4 jj
        {
            float a[4][4];
#pragma unroll
            for (int t = 0; t < 4; ++t) a[t][0] = a[t][1] = a[t][2] = a[t][3] = 0.f;
            const float* dz0 = dz_s + (4 * w) * ENF_H;
#pragma unroll 4
            for (int k = 0; k < ENF_H; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(wa_s + k * ENF_H + 4 * lane);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float d = dz0[t * ENF_H + k];
                    a[t][0] = fmaf(d, wv.x, a[t][0]); a[t][1] = fmaf(d, wv.y, a[t][1]);
                    a[t][2] = fmaf(d, wv.z, a[t][2]); a[t][3] = fmaf(d, wv.w, a[t][3]);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = t0 + 4 * w + t;
                if (i < N) *reinterpret_cast<float4*>(dagg + (int64_t)i * ENF_H + 4 * lane) = make_float4(a[t][0], a[t][1], a[t][2], a[t][3]);
            }
        }
        // ---- dW4[k][j] += sum_i dz[i][k] in[i][j]; thread = hidden 4lane..+3 x inputs 9w..9w+8
#pragma unroll 2
        for (int t = 0; t < BPT; ++t) {
            const float4 dv = *reinterpret_cast<const float4*>(dz_s + t * ENF_H + 4 * lane);
            const float* ip = in_s + t * BDP + JS * w;
#pragma unroll
            for (int jj = 0; jj < JS; ++jj) {
                const float a = ip[jj];
                wacc[0][jj] = fmaf(dv.x, a, wacc[0][jj]); wacc[1][jj] = fmaf(dv.y, a, wacc[1][jj]);
                wacc[2][jj] = fmaf(dv.z, a, wacc[2][jj]); wacc[3][jj] = fmaf(dv.w, a, wacc[3][jj]);
            }
        }
    }
    // ---- per-CTA partials
    float* p = partial + (int64_t)blockIdx.x * ((int64_t)ENF_H * D + ENF_H + nf * ENF_H + nf);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int jj = 0; jj < JS; ++jj) {
            const int j = JS * w + jj;
            if (j < D) p[(int64_t)(4 * lane + c) * D + j] = wacc[c][jj];
        }
    // db4 and dW5 rows: combine the 8 node groups (warps) in fixed order, one row at a time through smem
    p += (int64_t)ENF_H * D;
    float* red = dz_s;                 // [16 warps][128]
    const int R = 1 + nf;
    for (int r = 0; r < R; ++r) {
        __syncthreads();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            float v = gb4[c4];
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c)
                if (r == 1 + c) v = gw5[c][c4];
            red[w * ENF_H + 4 * lane + c4] = v;
        }
        __syncthreads();
        if (threadIdx.x < ENF_H) {
            float s = 0.f;
            for (int ww = 0; ww < BT / 32; ++ww) s += red[ww * ENF_H + threadIdx.x];
            p[r * ENF_H + threadIdx.x] = s;          // db4 [H] then dW5 [nf][H]: the partial layout order
        }
    }
    __syncthreads();
    if (lane < ENF_MAX_NF) red[w * ENF_MAX_NF + lane] = gb5;
    __syncthreads();
    if (threadIdx.x < nf) {
        float s5 = 0.f;
        for (int ww = 0; ww < BT / 32; ++ww) s5 += red[ww * ENF_MAX_NF + threadIdx.x];
        p[R * ENF_H + threadIdx.x] = s5;
    }
}

// grad += sum over CTAs of the per-CTA partials; block = 32 elements x 8 CTA groups (group y adds CTAs y, y+8, ...
// in order, then the eight group sums are added in order): deterministic and coalesced
__global__ void __launch_bounds__(256) k_node_post_reduce(const float* __restrict__ partial, int n_cta, int stride, int D,
                                                           int nf, int o_w4, int o_b4, int o_w5, int o_b5,
                                                           float* __restrict__ grad) {
    int idx;
    float acc;
    if (!enf_reduce_partials_32x8(partial, n_cta, stride, idx, acc)) return;
    const int s0 = ENF_H * D, s1 = s0 + ENF_H, s2 = s1 + nf * ENF_H;
    int dst;
    if (idx < s0) dst = o_w4 + idx;
    else if (idx < s1) dst = o_b4 + (idx - s0);
    else if (idx < s2) dst = o_w5 + (idx - s1);
    else dst = o_b5 + (idx - s2);
    grad[dst] += acc;
}

}  // namespace

int enf_node_post_fwd(const float* h, const float* agg, int N, int nf, const float* lp, const float* packed,
                      float* z4, float* G, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const PackOffsets p = enf_pack_offsets(nf);
    const size_t smem = sizeof(float) * (DP * ENF_H + PT * DP);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_node_post_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    int grid = (N + PT - 1) / PT;
    if (grid > 2 * enf_num_sms()) grid = 2 * enf_num_sms();
    enf_count_launch(), k_node_post_fwd<<<grid, TPB, smem, st>>>(h, agg, N, nf, packed + p.w4t, lp + o.off[P_B4],
                                                                 lp + o.off[P_W5], lp + o.off[P_B5], z4, G);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_node_post_reduce(const float* partial, int n_cta, int nf, float* lgrad, cudaStream_t st) {
    const EgclOffsets o = enf_egcl_offsets(nf);
    const int D = nf + ENF_H;
    const int stride = ENF_H * D + ENF_H + nf * ENF_H + nf;
    enf_count_launch(), k_node_post_reduce<<<(stride + 31) / 32, 256, 0, st>>>(
        partial, n_cta, stride, D, nf, (int)o.off[P_W4], (int)o.off[P_B4], (int)o.off[P_W5], (int)o.off[P_B5], lgrad);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

static int node_post_bwd_grid(int N) {
    const int tiles = (N + BPT - 1) / BPT;
    const int cap = enf_num_sms();
    return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

int64_t enf_node_post_partial_floats(int N, int nf) {
    return (int64_t)node_post_bwd_grid(N) * ((int64_t)ENF_H * (nf + ENF_H) + ENF_H + nf * ENF_H + nf);
}

int enf_node_post_bwd(const float* h, const float* agg, const float* z4, const float* dG, int N, int nf,
                      const float* lp, const float* packed, float* dagg, float* dh, float* lgrad, float* partial,
                      cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const PackOffsets p = enf_pack_offsets(nf);
    const int grid = node_post_bwd_grid(N);
    const size_t smem = sizeof(float) * (BPT * BDP + BPT * ENF_H + BPT * ENF_MAX_NF + ENF_H * ENF_H);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_node_post_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    enf_count_launch(), k_node_post_bwd<<<grid, BT, smem, st>>>(h, agg, z4, dG, N, nf, lp + o.off[P_W4], packed + p.w4a,
                                                              lp + o.off[P_W5], dagg, dh, partial);
    ENF_CHECK_LAUNCH();
    return enf_node_post_reduce(partial, grid, nf, lgrad, st);
}
