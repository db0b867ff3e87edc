// Node-level (per atom) parts of one EGCL (enflow/nn/egcl.py):
//   node_pre : P = A h + b1, S = B h   (edge_nn.0 split algebraically: W1 [h_i; h_j; r] = A h_i + B h_j + w_r r + b1,
//              egcl.py:21-22,57-59)  and  Q = W7 silu(W6 h + b6) + b7   (vel_scaling_nn, egcl.py:52-55,91)
//   node_post: G = W5 silu(W4 [h; agg] + b4) + b5                    (node_nn, egcl.py:27-30,65-69)
// plus their backward passes.  Weight gradients are accumulated per CTA over a fixed, strided
// node-tile schedule and combined by k_reduce_partials in CTA order: deterministic.
#include "common.cuh"

namespace {

constexpr int NT = 16;     // nodes per tile
constexpr int TPB = 128;   // one thread per hidden unit

// ------------------------------------------------------------------------------------------ node_pre forward
__global__ void __launch_bounds__(TPB) k_node_pre_fwd(const float* __restrict__ h, int N, int nf,
                                                       const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ W6, const float* __restrict__ b6,
                                                       const float* __restrict__ W7, const float* __restrict__ b7,
                                                       float* __restrict__ P, float* __restrict__ S,
                                                       float* __restrict__ Q) {
    __shared__ float hs[NT][ENF_MAX_NF];
    __shared__ float red[NT][4];
    const int k = threadIdx.x, lane = k & 31, wid = k >> 5;
    float wa[ENF_MAX_NF], wb[ENF_MAX_NF], w6[ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        wa[c] = c < nf ? W1[k * (2 * nf + 1) + c] : 0.f;
        wb[c] = c < nf ? W1[k * (2 * nf + 1) + nf + c] : 0.f;
        w6[c] = c < nf ? W6[k * nf + c] : 0.f;
    }
    const float bb1 = b1[k], bb6 = b6[k], w7 = W7[k];
    for (int t0 = blockIdx.x * NT; t0 < N; t0 += gridDim.x * NT) {
        __syncthreads();
        for (int idx = k; idx < NT * ENF_MAX_NF; idx += TPB) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            hs[t][c] = (t0 + t < N && c < nf) ? h[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < NT; ++t) {
            float p = bb1, s = 0.f, z = bb6;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) {
                const float hv = hs[t][c];
                p = fmaf(wa[c], hv, p);
                s = fmaf(wb[c], hv, s);
                z = fmaf(w6[c], hv, z);
            }
            if (t0 + t < N) {
                P[(int64_t)(t0 + t) * ENF_H + k] = p;
                S[(int64_t)(t0 + t) * ENF_H + k] = s;
            }
            const float q = warp_sum(w7 * siluf_(z));
            if (lane == 0) red[t][wid] = q;
        }
        __syncthreads();
        if (k < NT && t0 + k < N) Q[t0 + k] = b7[0] + ((red[k][0] + red[k][1]) + (red[k][2] + red[k][3]));
    }
}

// ------------------------------------------------------------------------------------------ node_pre backward
// partial layout per CTA (compact): dW1 [H*(2nf+1)] | db1 [H] | dW6 [H*nf] | db6 [H] | dW7 [H] | db7 [1]
__global__ void __launch_bounds__(TPB) k_node_pre_bwd(const float* __restrict__ h, int N, int nf,
                                                       const float* __restrict__ W1, const float* __restrict__ W6,
                                                       const float* __restrict__ b6, const float* __restrict__ W7,
                                                       const float* __restrict__ dP, const float* __restrict__ dS,
                                                       const float* __restrict__ dQ, float* __restrict__ dh,
                                                       float* __restrict__ partial) {
    __shared__ float hs[NT][ENF_MAX_NF];
    __shared__ float dqs[NT];
    __shared__ float red[NT][ENF_MAX_NF][4];
    const int k = threadIdx.x, lane = k & 31, wid = k >> 5;
    float wa[ENF_MAX_NF], wb[ENF_MAX_NF], w6[ENF_MAX_NF];
    float gwa[ENF_MAX_NF], gwb[ENF_MAX_NF], gw6[ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        wa[c] = c < nf ? W1[k * (2 * nf + 1) + c] : 0.f;
        wb[c] = c < nf ? W1[k * (2 * nf + 1) + nf + c] : 0.f;
        w6[c] = c < nf ? W6[k * nf + c] : 0.f;
        gwa[c] = gwb[c] = gw6[c] = 0.f;
    }
    const float bb6 = b6[k], w7 = W7[k];
    float gb1 = 0.f, gb6 = 0.f, gw7 = 0.f, gb7 = 0.f;
    for (int t0 = blockIdx.x * NT; t0 < N; t0 += gridDim.x * NT) {
        __syncthreads();
        for (int idx = k; idx < NT * ENF_MAX_NF; idx += TPB) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            hs[t][c] = (t0 + t < N && c < nf) ? h[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        if (k < NT) dqs[k] = (t0 + k < N) ? dQ[t0 + k] : 0.f;
        __syncthreads();
#pragma unroll 2
        for (int t = 0; t < NT; ++t) {
            const bool ok = t0 + t < N;
            const float dp = ok ? dP[(int64_t)(t0 + t) * ENF_H + k] : 0.f;
            const float ds = ok ? dS[(int64_t)(t0 + t) * ENF_H + k] : 0.f;
            float z = bb6;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) z = fmaf(w6[c], hs[t][c], z);
            const float dq = dqs[t];
            const float sg = sigmoidf_(z);
            const float x6 = z * sg;
            const float dz6 = dq * w7 * (sg * (1.0f + z * (1.0f - sg)));
            gb1 += dp; gb6 += dz6; gw7 = fmaf(dq, x6, gw7);
            if (k == 0) gb7 += dq;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) {
                const float hv = hs[t][c];
                gwa[c] = fmaf(dp, hv, gwa[c]);
                gwb[c] = fmaf(ds, hv, gwb[c]);
                gw6[c] = fmaf(dz6, hv, gw6[c]);
                const float v = warp_sum(fmaf(wa[c], dp, fmaf(wb[c], ds, w6[c] * dz6)));
                if (lane == 0) red[t][c][wid] = v;
            }
        }
        __syncthreads();
        for (int idx = k; idx < NT * ENF_MAX_NF; idx += TPB) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            if (t0 + t < N && c < nf)
                dh[(int64_t)(t0 + t) * nf + c] += (red[t][c][0] + red[t][c][1]) + (red[t][c][2] + red[t][c][3]);
        }
    }
    const int e1 = 2 * nf + 1;
    float* p = partial + (int64_t)blockIdx.x * (ENF_H * e1 + ENF_H + ENF_H * nf + ENF_H + ENF_H + 1);
    for (int c = 0; c < nf; ++c) { p[k * e1 + c] = gwa[c]; p[k * e1 + nf + c] = gwb[c]; }
    p[k * e1 + 2 * nf] = 0.f;     // w_r column: its gradient comes from the edge kernel
    p += ENF_H * e1;
    p[k] = gb1; p += ENF_H;
    for (int c = 0; c < nf; ++c) p[k * nf + c] = gw6[c];
    p += ENF_H * nf;
    p[k] = gb6; p += ENF_H;
    p[k] = gw7; p += ENF_H;
    if (k == 0) p[0] = gb7;
}

// ------------------------------------------------------------------------------------------ partial reduce
struct SegTable {
    int n;
    int src[8], dst[8], len[8];
};

__global__ void k_reduce_partials(const float* __restrict__ partial, int n_cta, int64_t stride, SegTable segs,
                                  float* __restrict__ grad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int seg = -1, local = 0;
    for (int s = 0; s < segs.n; ++s)
        if (idx >= segs.src[s] && idx < segs.src[s] + segs.len[s]) { seg = s; local = idx - segs.src[s]; }
    if (seg < 0) return;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;        // fixed order: four interleaved chains
    int c = 0;
    for (; c + 4 <= n_cta; c += 4) {
        a0 += partial[(int64_t)(c + 0) * stride + idx];
        a1 += partial[(int64_t)(c + 1) * stride + idx];
        a2 += partial[(int64_t)(c + 2) * stride + idx];
        a3 += partial[(int64_t)(c + 3) * stride + idx];
    }
    for (; c < n_cta; ++c) a0 += partial[(int64_t)c * stride + idx];
    const float acc = (a0 + a1) + (a2 + a3);
    grad[segs.dst[seg] + local] += acc;
}

__global__ void k_transpose_pack(const float* __restrict__ W2, const float* __restrict__ W3,
                                 const float* __restrict__ W4, int nf, float* __restrict__ W2T,
                                 float* __restrict__ W3T, float* __restrict__ W4T, float* __restrict__ W4A) {
    const int D = nf + ENF_H;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < ENF_H * ENF_H) {
        const int k = idx / ENF_H, n = idx % ENF_H;     // out[k][n] = W[n][k]
        W2T[idx] = W2[n * ENF_H + k];
        W3T[idx] = W3[n * ENF_H + k];
        W4A[idx] = W4[(int64_t)k * D + nf + n];         // W4A[k][jj] = W4[k][nf + jj] (aligned rows)
    }
    if (idx < D * ENF_H) {
        const int j = idx / ENF_H, k = idx % ENF_H;     // W4T[j][k] = W4[k][j]
        W4T[idx] = W4[(int64_t)k * D + j];
    }
}

}  // namespace

int enf_node_grid(int N) {
    int tiles = (N + NT - 1) / NT;
    int cap = enf_num_sms() * 4;
    return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

int enf_pack_layer(const float* layer_params, int nf, float* packed, cudaStream_t st) {
    const EgclOffsets o = enf_egcl_offsets(nf);
    const PackOffsets p = enf_pack_offsets(nf);
    const int total = (nf + ENF_H) * ENF_H;
    enf_count_launch(), k_transpose_pack<<<(total + 255) / 256, 256, 0, st>>>(layer_params + o.off[P_W2], layer_params + o.off[P_W3],
                                                          layer_params + o.off[P_W4], nf, packed + p.w2t,
                                                          packed + p.w3t, packed + p.w4t, packed + p.w4a);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_node_pre_fwd(const float* h, int N, int nf, const float* lp, float* P, float* S, float* Q, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    enf_count_launch(), k_node_pre_fwd<<<enf_node_grid(N), TPB, 0, st>>>(h, N, nf, lp + o.off[P_W1], lp + o.off[P_B1], lp + o.off[P_W6],
                                                     lp + o.off[P_B6], lp + o.off[P_W7], lp + o.off[P_B7], P, S, Q);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

static int node_pre_bwd_grid(int N) {      // fewer CTAs than the forward: every CTA leaves a partial to reduce
    const int g = enf_node_grid(N), cap = enf_num_sms();
    return g < cap ? g : cap;
}

int64_t enf_node_pre_partial_floats(int N, int nf) {
    return (int64_t)node_pre_bwd_grid(N) * (ENF_H * (2 * nf + 1) + ENF_H + ENF_H * nf + ENF_H + ENF_H + 1);
}

int enf_node_pre_bwd(const float* h, int N, int nf, const float* lp, const float* dP, const float* dS,
                     const float* dQ, float* dh, float* lgrad, float* partial, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const int grid = node_pre_bwd_grid(N);
    enf_count_launch(), k_node_pre_bwd<<<grid, TPB, 0, st>>>(h, N, nf, lp + o.off[P_W1], lp + o.off[P_W6], lp + o.off[P_B6],
                                         lp + o.off[P_W7], dP, dS, dQ, dh, partial);
    SegTable s;
    const int e1 = 2 * nf + 1;
    int src = 0;
    s.n = 6;
    const int lens[6] = {ENF_H * e1, ENF_H, ENF_H * nf, ENF_H, ENF_H, 1};
    const int dsts[6] = {(int)o.off[P_W1], (int)o.off[P_B1], (int)o.off[P_W6], (int)o.off[P_B6], (int)o.off[P_W7],
                         (int)o.off[P_B7]};
    for (int i = 0; i < 6; ++i) { s.src[i] = src; s.dst[i] = dsts[i]; s.len[i] = lens[i]; src += lens[i]; }
    enf_count_launch(), k_reduce_partials<<<(src + 255) / 256, 256, 0, st>>>(partial, grid, src, s, lgrad);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

