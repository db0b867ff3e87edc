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
// Per 32-node tile, 256 threads:
//   stage  : dP / dS rows -> shared memory (coalesced float4)
//   phase 1: thread = (hidden unit k, half of the tile): dz6 = dQ w7 silu'(z6) into shared memory, weight
//            gradients (dA, dB, dW6, db1, db6, dW7) accumulated in registers over the CTA's whole tile schedule
//   phase 2: thread = (node, input column c): dh[node][c] += A[:,c].dP + B[:,c].dS + W6[:,c].dz6 as three
//            128-long dot products from shared memory (no cross-thread reduction)
// partial layout per CTA (compact): dW1 [H*(2nf+1)] | db1 [H] | dW6 [H*nf] | db6 [H] | dW7 [H] | db7 [1]
constexpr int BN = 32;       // nodes per tile
constexpr int BTH = 256;
constexpr int VP = 132;      // padded row (floats): conflict-free float4 reads in phase 2
static_assert(ENF_MAX_NF == 8 && BN * ENF_MAX_NF == BTH && ENF_H == 128, "thread maps of k_node_pre_bwd");

__global__ void __launch_bounds__(BTH) k_node_pre_bwd(const float* __restrict__ h, int N, int nf,
                                                       const float* __restrict__ W1, const float* __restrict__ W6,
                                                       const float* __restrict__ b6, const float* __restrict__ W7,
                                                       const float* __restrict__ dP, const float* __restrict__ dS,
                                                       const float* __restrict__ dQ, float* __restrict__ dh,
                                                       float* __restrict__ partial) {
    extern __shared__ __align__(16) float smem_pre[];
    float* vP = smem_pre;                          // [BN][VP]
    float* vS = vP + BN * VP;
    float* vZ = vS + BN * VP;
    float* ws = vZ + BN * VP;                      // [3][ENF_MAX_NF][VP]: A^T, B^T, W6^T rows
    float* hs = ws + 3 * ENF_MAX_NF * VP;          // [BN][ENF_MAX_NF]
    float* dqs = hs + BN * ENF_MAX_NF;             // [BN]
    const int tid = threadIdx.x, k = tid & (ENF_H - 1), half = tid >> 7;
    const int e1 = 2 * nf + 1;
    for (int idx = tid; idx < 3 * ENF_MAX_NF * ENF_H; idx += BTH) {
        const int m = idx / (ENF_MAX_NF * ENF_H), c = (idx / ENF_H) % ENF_MAX_NF, kk = idx % ENF_H;
        float v = 0.f;
        if (c < nf) v = m == 0 ? W1[kk * e1 + c] : (m == 1 ? W1[kk * e1 + nf + c] : W6[kk * nf + c]);
        ws[(m * ENF_MAX_NF + c) * VP + kk] = v;
    }
    float w6[ENF_MAX_NF], gwa[ENF_MAX_NF], gwb[ENF_MAX_NF], gw6[ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        w6[c] = c < nf ? W6[k * nf + c] : 0.f;
        gwa[c] = gwb[c] = gw6[c] = 0.f;
    }
    const float bb6 = b6[k], w7 = W7[k];
    float gb1 = 0.f, gb6 = 0.f, gw7 = 0.f, gb7 = 0.f;
    const int pt = tid >> 3, pc = tid & 7;         // phase 2: node within the tile, input column
    for (int t0 = blockIdx.x * BN; t0 < N; t0 += gridDim.x * BN) {
        __syncthreads();
        for (int idx = tid; idx < BN * (ENF_H / 4); idx += BTH) {
            const int t = idx >> 5, k4 = idx & 31;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (t0 + t < N) {
                a = __ldg(reinterpret_cast<const float4*>(dP + (int64_t)(t0 + t) * ENF_H) + k4);
                b = __ldg(reinterpret_cast<const float4*>(dS + (int64_t)(t0 + t) * ENF_H) + k4);
            }
            *reinterpret_cast<float4*>(vP + t * VP + 4 * k4) = a;
            *reinterpret_cast<float4*>(vS + t * VP + 4 * k4) = b;
        }
        {
            const int t = tid >> 3, c = tid & 7;   // BN * ENF_MAX_NF == BTH
            hs[tid] = (t0 + t < N && c < nf) ? h[(int64_t)(t0 + t) * nf + c] : 0.f;
            if (tid < BN) dqs[tid] = (t0 + tid < N) ? dQ[t0 + tid] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int tt = 0; tt < BN / 2; ++tt) {
            const int t = half * (BN / 2) + tt;
            const float dp = vP[t * VP + k], ds = vS[t * VP + k];
            const float4 h0 = *reinterpret_cast<const float4*>(hs + t * ENF_MAX_NF);
            const float4 h1 = *reinterpret_cast<const float4*>(hs + t * ENF_MAX_NF + 4);
            const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            float z = bb6;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) z = fmaf(w6[c], hv[c], z);
            const float dq = dqs[t];
            const float sg = sigmoidf_(z);
            const float x6 = z * sg;
            const float dz6 = dq * w7 * (sg * (1.0f + z * (1.0f - sg)));
            vZ[t * VP + k] = dz6;
            gb1 += dp; gb6 += dz6; gw7 = fmaf(dq, x6, gw7);
            gb7 += dq;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) {
                gwa[c] = fmaf(dp, hv[c], gwa[c]);
                gwb[c] = fmaf(ds, hv[c], gwb[c]);
                gw6[c] = fmaf(dz6, hv[c], gw6[c]);
            }
        }
        __syncthreads();
        if (pc < nf && t0 + pt < N) {
            const float4* a = reinterpret_cast<const float4*>(ws + (0 * ENF_MAX_NF + pc) * VP);
            const float4* b = reinterpret_cast<const float4*>(ws + (1 * ENF_MAX_NF + pc) * VP);
            const float4* c6 = reinterpret_cast<const float4*>(ws + (2 * ENF_MAX_NF + pc) * VP);
            const float4* xp = reinterpret_cast<const float4*>(vP + pt * VP);
            const float4* xs = reinterpret_cast<const float4*>(vS + pt * VP);
            const float4* xz = reinterpret_cast<const float4*>(vZ + pt * VP);
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 4
            for (int k4 = 0; k4 < ENF_H / 4; ++k4) {
                const float4 wa = a[k4], wb = b[k4], wz = c6[k4], p4 = xp[k4], q4 = xs[k4], z4 = xz[k4];
                s0 = fmaf(wa.x, p4.x, fmaf(wa.y, p4.y, fmaf(wa.z, p4.z, fmaf(wa.w, p4.w, s0))));
                s1 = fmaf(wb.x, q4.x, fmaf(wb.y, q4.y, fmaf(wb.z, q4.z, fmaf(wb.w, q4.w, s1))));
                s2 = fmaf(wz.x, z4.x, fmaf(wz.y, z4.y, fmaf(wz.z, z4.z, fmaf(wz.w, z4.w, s2))));
            }
            dh[(int64_t)(t0 + pt) * nf + pc] += (s0 + s1) + s2;
        }
    }
    // ---- per-CTA partial: the two halves are combined in fixed order through shared memory
    __syncthreads();
    float* red = vP;                               // [2][27][128] floats <= 3 * BN * VP
    constexpr int NV = 3 * ENF_MAX_NF + 3;
    float* mine = red + half * NV * ENF_H;
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) {
        mine[(c) * ENF_H + k] = gwa[c];
        mine[(ENF_MAX_NF + c) * ENF_H + k] = gwb[c];
        mine[(2 * ENF_MAX_NF + c) * ENF_H + k] = gw6[c];
    }
    mine[(3 * ENF_MAX_NF + 0) * ENF_H + k] = gb1;
    mine[(3 * ENF_MAX_NF + 1) * ENF_H + k] = gb6;
    mine[(3 * ENF_MAX_NF + 2) * ENF_H + k] = gw7;
    __shared__ float gb7s[2];
    if (k == 0) gb7s[half] = gb7;
    __syncthreads();
    float* p = partial + (int64_t)blockIdx.x * (ENF_H * e1 + ENF_H + ENF_H * nf + ENF_H + ENF_H + 1);
    auto tot = [&](int v) { return red[v * ENF_H + k] + red[(NV + v) * ENF_H + k]; };
    if (half == 0) {
        for (int c = 0; c < nf; ++c) { p[k * e1 + c] = tot(c); p[k * e1 + nf + c] = tot(ENF_MAX_NF + c); }
        p[k * e1 + 2 * nf] = 0.f;     // w_r column: its gradient comes from the edge kernel
        p += ENF_H * e1;
        p[k] = tot(3 * ENF_MAX_NF); p += ENF_H;
        for (int c = 0; c < nf; ++c) p[k * nf + c] = tot(2 * ENF_MAX_NF + c);
        p += ENF_H * nf;
        p[k] = tot(3 * ENF_MAX_NF + 1); p += ENF_H;
        p[k] = tot(3 * ENF_MAX_NF + 2); p += ENF_H;
        if (k == 0) p[0] = gb7s[0] + gb7s[1];
    }
}

// ------------------------------------------------------------------------------------------ partial reduce
struct SegTable {
    int n;
    int src[8], dst[8], len[8];
};

// grad[dst(idx)] += sum over CTAs of partial[cta][idx].  Block = 32 elements x 8 CTA groups: group y adds
// CTAs y, y+8, ... in order, the eight group sums are then added in order: deterministic and coalesced.
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ partial, int n_cta, int64_t stride,
                                                          SegTable segs, float* __restrict__ grad) {
    int idx;
    float acc;
    if (!enf_reduce_partials_32x8(partial, n_cta, stride, idx, acc)) return;
    int seg = -1, local = 0;
    for (int s = 0; s < segs.n; ++s)
        if (idx >= segs.src[s] && idx < segs.src[s] + segs.len[s]) { seg = s; local = idx - segs.src[s]; }
    if (seg >= 0) grad[segs.dst[seg] + local] += acc;
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
    int cap = enf_num_sms() * 8;      // 32 warps per SM: the kernel is latency-bound (three barriers per 16 nodes)
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

static int node_pre_bwd_grid(int N) {      // persistent: every CTA leaves a partial to reduce
    const int g = (N + BN - 1) / BN, cap = 3 * enf_num_sms();
    return g < cap ? (g > 0 ? g : 1) : cap;
}

int64_t enf_node_pre_partial_floats(int N, int nf) {
    return (int64_t)node_pre_bwd_grid(N) * (ENF_H * (2 * nf + 1) + ENF_H + ENF_H * nf + ENF_H + ENF_H + 1);
}

int enf_node_pre_bwd(const float* h, int N, int nf, const float* lp, const float* dP, const float* dS,
                     const float* dQ, float* dh, float* lgrad, float* partial, cudaStream_t st, cudaStream_t st_red) {
    if (N == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    const int grid = node_pre_bwd_grid(N);
    const size_t smem = sizeof(float) * (3 * BN * VP + 3 * ENF_MAX_NF * VP + BN * ENF_MAX_NF + BN);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_node_pre_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    enf_count_launch(), k_node_pre_bwd<<<grid, BTH, smem, st>>>(h, N, nf, lp + o.off[P_W1], lp + o.off[P_W6], lp + o.off[P_B6],
                                         lp + o.off[P_W7], dP, dS, dQ, dh, partial);
    SegTable s;
    const int e1 = 2 * nf + 1;
    int src = 0;
    s.n = 6;
    const int lens[6] = {ENF_H * e1, ENF_H, ENF_H * nf, ENF_H, ENF_H, 1};
    const int dsts[6] = {(int)o.off[P_W1], (int)o.off[P_B1], (int)o.off[P_W6], (int)o.off[P_B6], (int)o.off[P_W7],
                         (int)o.off[P_B7]};
    for (int i = 0; i < 6; ++i) { s.src[i] = src; s.dst[i] = dsts[i]; s.len[i] = lens[i]; src += lens[i]; }
    enf_chain(st, st_red);
    enf_count_launch(), k_reduce_partials<<<(src + 31) / 32, 256, 0, st_red>>>(partial, grid, src, s, lgrad);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

