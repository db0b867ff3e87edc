// K4: variational argmax dequantiser (enflow/nn/argmax.py:14-26), forward and backward.
//   net = Wa2 silu(Wa0 h + ba0) + ba2 ; [log_s, t] = chunk(net) ; u = t + eps * exp(log_s)
//   T = sum_c h_c u_c ; z = h u + (1-h)(T - softplus(T-u))
//   log_q = -1/2 (sum u^2 + log 2pi) - sum log_s - sum (1-h) logsigmoid(T-u)
// The noise eps (float32 in the reference, argmax.py:17) is an input so that the oracle and this
// kernel consume identical draws.  log_q is returned as per-atom terms (the constant -1/2 log 2pi is
// added once by the caller, helpers.py:4-5 / quirk Q3); they are reduced per molecule in fixed order.
#include "common.cuh"

namespace {

constexpr int NT = 16;
constexpr int TPB = 128;

__device__ __forceinline__ float softplusf_(float x) {   // F.softplus, threshold 20
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float logsigmoidf_(float x) {   // min(x,0) - log1p(exp(-|x|))
    return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}

// net[t][o] = b2[o] + sum_k W2[o][k] x0[t][k]; warp w handles nodes w, w+4, ...
__device__ __forceinline__ void head(const float* __restrict__ W2, const float* __restrict__ b2, int nf,
                                     const float* x0, float* net) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int t = wid; t < NT; t += 4) {
        for (int o = 0; o < 2 * nf; ++o) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) s = fmaf(W2[o * ENF_H + lane + 32 * q], x0[t * ENF_H + lane + 32 * q], s);
            s = warp_sum(s);
            if (lane == 0) net[t * 2 * ENF_MAX_NF + o] = s + b2[o];
        }
    }
}

__global__ void __launch_bounds__(TPB) k_argmax_fwd(const float* __restrict__ h, const float* __restrict__ eps, int N,
                                                     int nf, const float* __restrict__ W0, const float* __restrict__ b0,
                                                     const float* __restrict__ W2, const float* __restrict__ b2,
                                                     float* __restrict__ z, float* __restrict__ logq_atom) {
    __shared__ float hs[NT][ENF_MAX_NF];
    __shared__ float x0[NT * ENF_H];
    __shared__ float net[NT * 2 * ENF_MAX_NF];
    const int k = threadIdx.x;
    float w0[ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) w0[c] = c < nf ? W0[k * nf + c] : 0.f;
    const float bb0 = b0[k];
    for (int t0 = blockIdx.x * NT; t0 < N; t0 += gridDim.x * NT) {
        __syncthreads();
        for (int idx = k; idx < NT * ENF_MAX_NF; idx += TPB) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            hs[t][c] = (t0 + t < N && c < nf) ? h[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < NT; ++t) {
            float a = bb0;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) a = fmaf(w0[c], hs[t][c], a);
            x0[t * ENF_H + k] = siluf_(a);
        }
        __syncthreads();
        head(W2, b2, nf, x0, net);
        __syncthreads();
        if (k < NT && t0 + k < N) {
            const int i = t0 + k;
            float u[ENF_MAX_NF];
            float T = 0.f, lq = 0.f;
            for (int c = 0; c < nf; ++c) {
                const float ls = net[k * 2 * ENF_MAX_NF + c], tr = net[k * 2 * ENF_MAX_NF + nf + c];
                u[c] = fmaf(eps[(int64_t)i * nf + c], expf(ls), tr);
                T = fmaf(hs[k][c], u[c], T);
                lq -= 0.5f * u[c] * u[c] + ls;
            }
            for (int c = 0; c < nf; ++c) {
                const float hv = hs[k][c], a = T - u[c];
                z[(int64_t)i * nf + c] = hv * u[c] + (1.f - hv) * (T - softplusf_(a));
                lq -= (1.f - hv) * logsigmoidf_(a);
            }
            logq_atom[i] = lq;
        }
    }
}

// partial layout per CTA: dW0 [H*nf] | db0 [H] | dW2 [2nf*H] | db2 [2nf]
__global__ void __launch_bounds__(TPB) k_argmax_bwd(const float* __restrict__ h, const float* __restrict__ eps, int N,
                                                     int nf, const float* __restrict__ W0, const float* __restrict__ b0,
                                                     const float* __restrict__ W2, const float* __restrict__ b2,
                                                     const float* __restrict__ dz, const float* __restrict__ dlogq,
                                                     float* __restrict__ partial) {
    __shared__ float hs[NT][ENF_MAX_NF];
    __shared__ float x0[NT * ENF_H];
    __shared__ float net[NT * 2 * ENF_MAX_NF];
    __shared__ float dnet[NT * 2 * ENF_MAX_NF];
    const int k = threadIdx.x;
    float w0[ENF_MAX_NF], gw0[ENF_MAX_NF], w2[2 * ENF_MAX_NF], gw2[2 * ENF_MAX_NF];
#pragma unroll
    for (int c = 0; c < ENF_MAX_NF; ++c) { w0[c] = c < nf ? W0[k * nf + c] : 0.f; gw0[c] = 0.f; }
#pragma unroll
    for (int o = 0; o < 2 * ENF_MAX_NF; ++o) { w2[o] = o < 2 * nf ? W2[o * ENF_H + k] : 0.f; gw2[o] = 0.f; }
    const float bb0 = b0[k];
    const float dl = dlogq[0];
    float gb0 = 0.f, gb2 = 0.f;
    for (int t0 = blockIdx.x * NT; t0 < N; t0 += gridDim.x * NT) {
        __syncthreads();
        for (int idx = k; idx < NT * ENF_MAX_NF; idx += TPB) {
            const int t = idx / ENF_MAX_NF, c = idx % ENF_MAX_NF;
            hs[t][c] = (t0 + t < N && c < nf) ? h[(int64_t)(t0 + t) * nf + c] : 0.f;
        }
        __syncthreads();
        float z0[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float a = bb0;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) a = fmaf(w0[c], hs[t][c], a);
            z0[t] = a;
            x0[t * ENF_H + k] = siluf_(a);
        }
        __syncthreads();
        head(W2, b2, nf, x0, net);
        __syncthreads();
        if (k < NT) {
            const int i = t0 + k;
            for (int o = 0; o < 2 * ENF_MAX_NF; ++o) dnet[k * 2 * ENF_MAX_NF + o] = 0.f;
            if (i < N) {
                float u[ENF_MAX_NF], sg[ENF_MAX_NF], es[ENF_MAX_NF];
                float T = 0.f;
                for (int c = 0; c < nf; ++c) {
                    const float ls = net[k * 2 * ENF_MAX_NF + c], tr = net[k * 2 * ENF_MAX_NF + nf + c];
                    es[c] = eps[(int64_t)i * nf + c] * expf(ls);
                    u[c] = es[c] + tr;
                    T = fmaf(hs[k][c], u[c], T);
                }
                float gT = 0.f;
                for (int c = 0; c < nf; ++c) {
                    sg[c] = sigmoidf_(T - u[c]);
                    const float w = (1.f - hs[k][c]) * (1.f - sg[c]);
                    gT += (dz[(int64_t)i * nf + c] - dl) * w;
                }
                for (int c = 0; c < nf; ++c) {
                    const float hv = hs[k][c], dzc = dz[(int64_t)i * nf + c];
                    float gu = dzc * (hv + (1.f - hv) * sg[c]) + dl * (-u[c] + (1.f - hv) * (1.f - sg[c])) + gT * hv;
                    dnet[k * 2 * ENF_MAX_NF + nf + c] = gu;                 // d translate
                    dnet[k * 2 * ENF_MAX_NF + c] = gu * es[c] - dl;          // d log_scale
                }
            }
        }
        __syncthreads();
        if (k < 2 * nf)
            for (int t = 0; t < NT; ++t) gb2 += dnet[t * 2 * ENF_MAX_NF + k];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float dx = 0.f;
            const float x = x0[t * ENF_H + k];
#pragma unroll
            for (int o = 0; o < 2 * ENF_MAX_NF; ++o) {
                const float d = dnet[t * 2 * ENF_MAX_NF + o];
                dx = fmaf(w2[o], d, dx);
                gw2[o] = fmaf(d, x, gw2[o]);
            }
            const float d0 = dx * dsiluf_(z0[t]);
            gb0 += d0;
#pragma unroll
            for (int c = 0; c < ENF_MAX_NF; ++c) gw0[c] = fmaf(d0, hs[t][c], gw0[c]);
        }
    }
    float* p = partial + (int64_t)blockIdx.x * (ENF_H * nf + ENF_H + 2 * nf * ENF_H + 2 * nf);
    for (int c = 0; c < nf; ++c) p[k * nf + c] = gw0[c];
    p += ENF_H * nf;
    p[k] = gb0; p += ENF_H;
    for (int o = 0; o < 2 * nf; ++o) p[o * ENF_H + k] = gw2[o];
    p += 2 * nf * ENF_H;
    if (k < 2 * nf) p[k] = gb2;
}

__global__ void k_argmax_reduce(const float* __restrict__ partial, int n_cta, int stride, int nf,
                                int o_w0, int o_b0, int o_w2, int o_b2, float* __restrict__ grad) {
    int idx;
    float acc;
    if (!enf_reduce_partials_32x8(partial, n_cta, stride, idx, acc)) return;
    int dst;
    const int s0 = ENF_H * nf, s1 = s0 + ENF_H, s2 = s1 + 2 * nf * ENF_H;
    if (idx < s0) dst = o_w0 + idx;
    else if (idx < s1) dst = o_b0 + (idx - s0);
    else if (idx < s2) dst = o_w2 + (idx - s1);
    else dst = o_b2 + (idx - s2);
    grad[dst] += acc;
}

// per-molecule fixed-order sum of a per-atom float quantity into a double
__global__ void __launch_bounds__(256) k_mol_sum(const float* __restrict__ x, const int* __restrict__ mol_off, int B,
                                                  double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < B; m += warps) {
        double s = 0.0;
        for (int i = mol_off[m] + lane; i < mol_off[m + 1]; i += 32) s += (double)x[i];
        s = warp_sum(s);
        if (lane == 0) out[m] = s;
    }
}

// log_q = sum_m logq_mol[m] - 1/2 log(2 pi)    (single CTA, fixed order)
__global__ void __launch_bounds__(256) k_total(const double* __restrict__ x, int B, double add, float* __restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) s += x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = (float)(t + add);
    }
}

// ArgMax.reverse (argmax.py:28-29): one_hot(argmax(z)); first maximum wins like torch.argmax
__global__ void k_argmax_reverse(float* __restrict__ h, int N, int nf) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        int best = 0;
        float bv = h[(int64_t)i * nf];
        for (int c = 1; c < nf; ++c) {
            const float v = h[(int64_t)i * nf + c];
            if (v > bv) { bv = v; best = c; }
        }
        for (int c = 0; c < nf; ++c) h[(int64_t)i * nf + c] = c == best ? 1.f : 0.f;
    }
}

}  // namespace

int enf_argmax_reverse(float* h, int N, int nf, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    int blocks = (N + 255) / 256;
    if (blocks > enf_num_sms() * 8) blocks = enf_num_sms() * 8;
    enf_count_launch(), k_argmax_reverse<<<blocks, 256, 0, st>>>(h, N, nf);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

static int tile_grid(int N) {
    int tiles = (N + NT - 1) / NT;
    int cap = enf_num_sms() * 4;
    return tiles < cap ? (tiles > 0 ? tiles : 1) : cap;
}

int enf_argmax_fwd(const float* h, const float* eps, int N, int nf, const float* ap, const int* mol_off, int B,
                   float* z, float* logq_atom, double* logq_mol, float* log_q, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const ArgmaxOffsets o = enf_argmax_offsets(nf);
    enf_count_launch(), k_argmax_fwd<<<tile_grid(N), TPB, 0, st>>>(h, eps, N, nf, ap + o.off[PA_W0], ap + o.off[PA_B0], ap + o.off[PA_W2],
                                               ap + o.off[PA_B2], z, logq_atom);
    int mg = (B + 7) / 8;
    if (mg > enf_num_sms() * 8) mg = enf_num_sms() * 8;
    enf_count_launch(), k_mol_sum<<<mg, 256, 0, st>>>(logq_atom, mol_off, B, logq_mol);
    enf_count_launch(), k_total<<<1, 256, 0, st>>>(logq_mol, B, -0.5 * 1.8378770664093453 /* log(2 pi) */, log_q);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int64_t enf_argmax_partial_floats(int N, int nf) {
    return (int64_t)tile_grid(N) * (ENF_H * nf + ENF_H + 2 * nf * ENF_H + 2 * nf);
}

int enf_argmax_bwd(const float* h, const float* eps, int N, int nf, const float* ap, const float* dz,
                   const float* dlogq, float* agrad, float* partial, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const ArgmaxOffsets o = enf_argmax_offsets(nf);
    const int grid = tile_grid(N);
    enf_count_launch(), k_argmax_bwd<<<grid, TPB, 0, st>>>(h, eps, N, nf, ap + o.off[PA_W0], ap + o.off[PA_B0], ap + o.off[PA_W2],
                                       ap + o.off[PA_B2], dz, dlogq, partial);
    const int stride = ENF_H * nf + ENF_H + 2 * nf * ENF_H + 2 * nf;
    enf_count_launch(), k_argmax_reduce<<<(stride + 31) / 32, 256, 0, st>>>(partial, grid, stride, nf, (int)o.off[PA_W0],
                                                           (int)o.off[PA_B0], (int)o.off[PA_W2], (int)o.off[PA_B2], agrad);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
