// K5: Alchemical_NLL (enflow/flow/loss.py:11-25), forward and backward.
//   per molecule: soft Lennard-Jones over all unordered pairs (no box, no cut-off; pairs with
//   r^2 == 0 are dropped, quirk Q13): 4 (1/s^6 - 1/s^3), s = r^2 + softening     (loss.py:14-18)
//   H = LJ + 1/2 sum vel^2 ; log_px = -H/kBT + logZ + ldj + log_gaussian(h) + log_gaussian(g) ; loss = -log_px / B
// One CTA per (molecule, slice of 128 atoms) with the molecule's positions staged in shared memory: thread = atom i, loop over
// all partners j (a 500-atom molecule is four CTAs; one CTA per molecule left 32 CTAs on 148 SMs with four atoms per
// thread).  Pair terms accumulate in fp64 per thread and are combined in a fixed order; mol_term holds one partial per
// (molecule, slice) and the batch scalar is finished by a single-CTA kernel that adds them in index order.
#include "common.cuh"

namespace {

constexpr int TPB = 128;

__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
}

// mol_term[m * S + s] = LJ_(m, atoms of slice s)/kBT  (+ (1/2 sum vel^2)/kBT + 1/2 sum h^2 + 1/2 sum g^2 of the molecule for s = 0)
// BWD: writes d loss / d(pos, vel, h, g) of the slice's atoms given scale = dloss / B.
template <bool BWD>
__global__ void __launch_bounds__(TPB) k_nll(const float* __restrict__ pos, const float* __restrict__ vel,
                                              const float* __restrict__ h, const float* __restrict__ g,
                                              const int* __restrict__ mol_off, int nf, float kBT, float softening,
                                              double* __restrict__ mol_term, const float* __restrict__ dloss,
                                              float inv_B, float* __restrict__ dpos, float* __restrict__ dvel,
                                              float* __restrict__ dh, float* __restrict__ dg) {
    extern __shared__ float ps[];     // [n][3]
    __shared__ double red[4];
    const int m = blockIdx.x, slice = blockIdx.y, S = gridDim.y;
    const int a0 = mol_off[m], n = mol_off[m + 1] - a0;
    const int i0 = slice * TPB, i1 = min(n, i0 + TPB);           // this CTA's atoms
    if (i0 >= n) {                                               // a molecule shorter than the longest one
        if (!BWD && threadIdx.x == 0) mol_term[(int64_t)m * S + slice] = 0.0;
        return;
    }
    for (int idx = threadIdx.x; idx < 3 * n; idx += TPB) ps[idx] = pos[(int64_t)a0 * 3 + idx];
    __syncthreads();
    const float scale = BWD ? dloss[0] * inv_B : 0.f;
    double acc = 0.0;
    for (int i = i0 + threadIdx.x; i < i1; i += TPB) {
        const float xi = ps[3 * i], yi = ps[3 * i + 1], zi = ps[3 * i + 2];
        double e = 0.0;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int j = 0; j < n; ++j) {
            const float dx = xi - ps[3 * j], dy = yi - ps[3 * j + 1], dz = zi - ps[3 * j + 2];
            const float r2 = dx * dx + dy * dy + dz * dz;
            if (j != i && r2 != 0.f) {
                const float s = r2 + softening;
                const float i3 = 1.0f / (s * s * s);
                e += (double)(4.0f * (i3 * i3 - i3));
                if (BWD) {
                    // d/ds 4(s^-6 - s^-3) = 4(-6 s^-7 + 3 s^-4); ds/dp_i = 2 (p_i - p_j)
                    const float de = 8.0f * (3.0f * i3 - 6.0f * i3 * i3) / s;
                    fx = fmaf(de, dx, fx); fy = fmaf(de, dy, fy); fz = fmaf(de, dz, fz);
                }
            }
        }
        acc += 0.5 * e;      // every unordered pair visited twice
        if (BWD) {
            const float c = scale / kBT;
            dpos[(int64_t)(a0 + i) * 3 + 0] = c * fx;
            dpos[(int64_t)(a0 + i) * 3 + 1] = c * fy;
            dpos[(int64_t)(a0 + i) * 3 + 2] = c * fz;
        }
    }
    if (BWD) {
        for (int idx = 3 * i0 + threadIdx.x; idx < 3 * i1; idx += TPB)
            dvel[(int64_t)a0 * 3 + idx] = scale / kBT * vel[(int64_t)a0 * 3 + idx];
        for (int idx = nf * i0 + threadIdx.x; idx < nf * i1; idx += TPB) {
            dh[(int64_t)a0 * nf + idx] = scale * h[(int64_t)a0 * nf + idx];
            dg[(int64_t)a0 * nf + idx] = scale * g[(int64_t)a0 * nf + idx];
        }
        return;
    }
    double ke = 0.0, gs = 0.0;
    if (slice == 0) {                                            // the molecule's Gaussian / kinetic terms: once
        for (int idx = threadIdx.x; idx < 3 * n; idx += TPB) { const double v = vel[(int64_t)a0 * 3 + idx]; ke += v * v; }
        for (int idx = threadIdx.x; idx < nf * n; idx += TPB) {
            const double a = h[(int64_t)a0 * nf + idx], b = g[(int64_t)a0 * nf + idx];
            gs += a * a + b * b;
        }
    }
    const double lj = block_sum_d(acc, red);
    const double k2 = block_sum_d(ke, red);
    const double g2 = block_sum_d(gs, red);
    if (threadIdx.x == 0) mol_term[(int64_t)m * S + slice] = (lj + 0.5 * k2) / (double)kBT + 0.5 * g2;
}

// loss = ( sum_m mol_term[m] - ldj - logZ + log(2 pi) ) / B       (loss.py:22-25 with helpers.py:4-5)
__global__ void __launch_bounds__(256) k_loss(const double* __restrict__ mol_term, int n_terms, int B,
                                               const float* __restrict__ ldj, double logZ, float* __restrict__ loss) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n_terms; i += 256) s += mol_term[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        loss[0] = (float)((t - (double)ldj[0] - logZ + 1.8378770664093453) / (double)B);
    }
}

// ldj = log_q + sum_m ldj_mol[m]  (dynamics.py:11,21), single CTA fixed order
__global__ void __launch_bounds__(256) k_ldj_total(const float* __restrict__ ldj_mol, int B,
                                                    const float* __restrict__ log_q, float* __restrict__ ldj) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) s += (double)ldj_mol[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        ldj[0] = (float)(t + (log_q ? (double)log_q[0] : 0.0));
    }
}

__global__ void k_neg_scale(const float* __restrict__ dloss, float inv_B, float* __restrict__ dldj) {
    dldj[0] = -dloss[0] * inv_B;
}

}  // namespace

#include <math.h>

// slices of 128 atoms per molecule: mol_term holds B * enf_nll_slices(max_n) doubles
int enf_nll_slices(int max_n) { return max_n > TPB ? (max_n + TPB - 1) / TPB : 1; }

int enf_nll_fwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B, int N,
                int nf, int max_n, float kBT, float softening, float z_lj, const float* ldj, double* mol_term,
                float* loss, cudaStream_t st) {
    if (B == 0) return ENF_OK;
    const size_t smem = sizeof(float) * 3 * (size_t)max_n;
    ENF_CHECK_ARG(smem <= 200 * 1024, "nll: molecule with %d atoms does not fit shared memory", max_n);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_nll<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int S = enf_nll_slices(max_n);
    enf_count_launch(), k_nll<false><<<dim3(B, S), TPB, smem, st>>>(pos, vel, h, g, mol_off, nf, kBT, softening, mol_term, nullptr, 0.f,
                                                nullptr, nullptr, nullptr, nullptr);
    const double logZ = -(double)N * (log((double)z_lj) - 1.5 * log(2.0 * M_PI / (double)kBT));
    enf_count_launch(), k_loss<<<1, 256, 0, st>>>(mol_term, B * S, B, ldj, logZ, loss);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_nll_bwd(const float* pos, const float* vel, const float* h, const float* g, const int* mol_off, int B, int nf,
                int max_n, float kBT, float softening, const float* dloss, float* dpos, float* dvel, float* dh,
                float* dg, float* dldj, cudaStream_t st) {
    if (B == 0) return ENF_OK;
    const size_t smem = sizeof(float) * 3 * (size_t)max_n;
    ENF_CHECK_ARG(smem <= 200 * 1024, "nll: molecule with %d atoms does not fit shared memory", max_n);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_nll<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    enf_count_launch(), k_nll<true><<<dim3(B, enf_nll_slices(max_n)), TPB, smem, st>>>(pos, vel, h, g, mol_off, nf, kBT, softening, nullptr,
                                                                   dloss, 1.0f / (float)B, dpos, dvel, dh, dg);
    enf_count_launch(), k_neg_scale<<<1, 1, 0, st>>>(dloss, 1.0f / (float)B, dldj);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_ldj_total(const float* ldj_mol, int B, const float* log_q, float* ldj, cudaStream_t st) {
    enf_count_launch(), k_ldj_total<<<1, 256, 0, st>>>(ldj_mol, B, log_q, ldj);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
