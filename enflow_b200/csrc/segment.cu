// K2: deterministic segmented reductions over row-sorted (CSR) edge data.
// Replaces unsorted_segment_sum / unsorted_segment_mean (enflow/utils/helpers.py:54-70), whose
// scatter_add_ is an unordered atomicAdd on CUDA.  Here every output row is owned by one warp that
// walks its segment in edge order, so the sum order is fixed (run-to-run bit-identical).
//
// Bandwidth shape: each lane owns 4 consecutive features (one 16 B load per edge row, 512 B per
// warp-wide request, fully coalesced); UNROLL independent loads are in flight per lane.
// Algorithmic bytes per call: E*H*4 (messages) + (N+1)*4 (rowptr) + N*H*4 (output).
#include "common.cuh"

namespace {

// 3-vector segment sum of one row by one warp (lanes stride over the row's edges, xor tree): the k_segment_sum3
// body, callable from the 128-wide kernels so a row's two reductions share one launch
template <bool PERM>
__device__ __forceinline__ void row_sum3(const float* __restrict__ x, const int* __restrict__ perm, int e0, int e1, int i,
                                         int lane, int mean, float scale, int accumulate, float* __restrict__ out) {
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int e = e0 + lane; e < e1; e += 32) {
        const int src = PERM ? perm[e] : e;
        ax += x[(int64_t)src * 3 + 0];
        ay += x[(int64_t)src * 3 + 1];
        az += x[(int64_t)src * 3 + 2];
    }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane < 3) {
        float v = lane == 0 ? ax : (lane == 1 ? ay : az);
        const int deg = e1 - e0;
        if (mean) v = v / (float)(deg > 1 ? deg : 1);
        v *= scale;
        float* o = out + (int64_t)i * 3 + lane;
        *o = accumulate ? (*o + v) : v;
    }
}

template <bool SILU, bool PERM>
__global__ void __launch_bounds__(256) k_segment_sum128(const float* __restrict__ x, const int* __restrict__ ptr,
                                                         const int* __restrict__ perm, int N, int E_cap,
                                                         float* __restrict__ out, const float* __restrict__ x3,
                                                         float scale3, float* __restrict__ out3,
                                                         const int* __restrict__ rowptr3) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int ii = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ii < N; ii += warps_per_grid) {
        // Segments are walked from the LAST one: `x` was written front to back by the kernel before this one and is
        // larger than the L2 (E*H*4 = 426 MB on C2), so its tail is what is still resident when this kernel starts.
        const int i = N - 1 - ii;
        int e0 = ptr[i], e1 = ptr[i + 1];
        if (e1 > E_cap) e1 = E_cap;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = e0;
        constexpr int UNROLL = 4;
        for (; e + UNROLL <= e1; e += UNROLL) {
            float4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int src = PERM ? perm[e + u] : (e + u);
                v[u] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)src * ENF_H) + lane);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (SILU) { v[u].x = siluf_(v[u].x); v[u].y = siluf_(v[u].y); v[u].z = siluf_(v[u].z); v[u].w = siluf_(v[u].w); }
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
        }
        for (; e < e1; ++e) {
            const int src = PERM ? perm[e] : e;
            float4 v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)src * ENF_H) + lane);
            if (SILU) { v.x = siluf_(v.x); v.y = siluf_(v.y); v.z = siluf_(v.z); v.w = siluf_(v.w); }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out + (int64_t)i * ENF_H)[lane] = acc;
        if (x3) {
            // rowptr3: first out3 += the DIRECT sum of x3 over row i's own edges, then the (permuted) segment's share:
            // both halves of d loss / d pos (data/base.py:17) in one launch, in the order two launches used to add them
            if (rowptr3) {
                const int r0 = rowptr3[i];
                int r1 = rowptr3[i + 1];
                if (r1 > E_cap) r1 = E_cap;
                row_sum3<false>(x3, nullptr, r0, r1, i, lane, 0, 1.0f, 1, out3);
                __syncwarp();
            }
            row_sum3<PERM>(x3, perm, e0, e1, i, lane, 0, scale3, 1, out3);      // out3 += scale3 * sum
        }
    }
}

// 3-vector per edge (coordinate updates): lanes stride over the segment, then a fixed xor tree.
// out[i] = scale * sum / (MEAN ? max(deg,1) : 1)   (helpers.py:70 count.clamp(min=1), quirk Q12)
// With sign/accumulate options it also serves the backward scatter of d(coord_diff) onto positions.
template <bool PERM>
__global__ void __launch_bounds__(256) k_segment_sum3(const float* __restrict__ x, const int* __restrict__ ptr,
                                                       const int* __restrict__ perm, int N, int E_cap, int mean,
                                                       float scale, int accumulate, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps_per_grid) {
        int e0 = ptr[i], e1 = ptr[i + 1];
        if (e1 > E_cap) e1 = E_cap;
        float ax = 0.f, ay = 0.f, az = 0.f;
        for (int e = e0 + lane; e < e1; e += 32) {
            const int src = PERM ? perm[e] : e;
            ax += x[(int64_t)src * 3 + 0];
            ay += x[(int64_t)src * 3 + 1];
            az += x[(int64_t)src * 3 + 2];
        }
        ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
        if (lane < 3) {
            float v = lane == 0 ? ax : (lane == 1 ? ay : az);
            const int deg = e1 - e0;
            if (mean) v = v / (float)(deg > 1 ? deg : 1);
            v *= scale;
            float* o = out + (int64_t)i * 3 + lane;
            *o = accumulate ? (*o + v) : v;
        }
    }
}

// ---- "runs": partial segment sums produced inside the tensor-core edge kernels ---------------------------
// A run is a maximal stretch of consecutive CSR edges that share a row AND a 16-edge block (the stretch one
// thread of an edge kernel owns).  run id of edge e = e/16 + #(rows <= row[e] whose first edge is not on a
// 16-boundary) = e/16 + mis[row[e]+1] with mis the exclusive scan of the flags below.  The edge kernels write
// one 128-wide partial per run (coalesced); k_run_sum128 adds the runs of each row in order (deterministic).
__global__ void k_run_flags(const int* __restrict__ rowptr, int N, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flags[i] = (rowptr[i + 1] > rowptr[i]) && (rowptr[i] & 15);
    if (i == N) flags[N] = 0;
}

__global__ void __launch_bounds__(256) k_run_sum128(const float* __restrict__ runs, const int* __restrict__ rowptr,
                                                     const int* __restrict__ mis, int N, int E_cap,
                                                     float* __restrict__ out, const float* __restrict__ x3, int mean3,
                                                     float scale3, int accumulate3, float* __restrict__ out3) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps_per_grid) {
        int e0 = rowptr[i], e1 = rowptr[i + 1];
        if (e1 > E_cap) e1 = E_cap;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e1 > e0) {
            const int m = mis[i + 1];
            const int r0 = (e0 >> 4) + m, r1 = ((e1 - 1) >> 4) + m;
            for (int r = r0; r <= r1; ++r) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(runs + (int64_t)r * ENF_H) + lane);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(out + (int64_t)i * ENF_H)[lane] = acc;
        if (x3) row_sum3<false>(x3, nullptr, e0, e1, i, lane, mean3, scale3, accumulate3, out3);
    }
}

}  // namespace

int enf_scan_int(int* a, int64_t n, int* sums, cudaStream_t st);

// mis: N+2 ints (exclusive scan of the misaligned-start flags, total at [N+1]); scratch: enf_scan_scratch_ints(N+1)
int enf_run_index(const int* rowptr, int N, int* mis, int* scratch, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    enf_count_launch(), k_run_flags<<<(N + 1 + 255) / 256, 256, 0, st>>>(rowptr, N, mis);
    ENF_CHECK_LAUNCH();
    return enf_scan_int(mis, N + 1, scratch, st);
}

int enf_run_sum128_sum3(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out,
                        const float* x3, int mean3, float scale3, int accumulate3, float* out3, cudaStream_t st);
int enf_segment_sum128_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu, float* out,
                            const float* x3, float scale3, float* out3, const int* rowptr3, cudaStream_t st);

int enf_run_sum128(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out, cudaStream_t st) {
    return enf_run_sum128_sum3(runs, rowptr, mis, N, E_cap, out, nullptr, 0, 0.f, 0, nullptr, st);
}

// the row's 128-wide run sum and (x3 != NULL) its 3-vector segment sum/mean in one launch
int enf_run_sum128_sum3(const float* runs, const int* rowptr, const int* mis, int N, int E_cap, float* out,
                        const float* x3, int mean3, float scale3, int accumulate3, float* out3, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const int blocks = min((N + 7) / 8, enf_num_sms() * 8);
    enf_count_launch(), k_run_sum128<<<blocks, 256, 0, st>>>(runs, rowptr, mis, N, E_cap, out, x3, mean3, scale3, accumulate3, out3);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_segment_sum128(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu, float* out,
                       cudaStream_t st) {
    return enf_segment_sum128_sum3(x, ptr, perm, N, E_cap, apply_silu, out, nullptr, 0.f, nullptr, nullptr, st);
}

// the 128-wide segment sum and (x3 != NULL) out3 += scale3 * the 3-vector segment sum over the same segments;
// rowptr3 != NULL: before that, out3 += the direct sum of x3 over the CSR row of the same index
int enf_segment_sum128_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int apply_silu, float* out,
                            const float* x3, float scale3, float* out3, const int* rowptr3, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const int blocks = min((N + 7) / 8, enf_num_sms() * 8);
    if (perm) {
        if (apply_silu) enf_count_launch(), k_segment_sum128<true, true><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, out, x3, scale3, out3, rowptr3);
        else enf_count_launch(), k_segment_sum128<false, true><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, out, x3, scale3, out3, rowptr3);
    } else {
        if (apply_silu) enf_count_launch(), k_segment_sum128<true, false><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, out, x3, scale3, out3, rowptr3);
        else enf_count_launch(), k_segment_sum128<false, false><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, out, x3, scale3, out3, rowptr3);
    }
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

int enf_segment_sum3(const float* x, const int* ptr, const int* perm, int N, int E_cap, int mean, float scale,
                     int accumulate, float* out, cudaStream_t st) {
    if (N == 0) return ENF_OK;
    const int blocks = min((N + 7) / 8, enf_num_sms() * 8);
    if (perm) enf_count_launch(), k_segment_sum3<true><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, mean, scale, accumulate, out);
    else enf_count_launch(), k_segment_sum3<false><<<blocks, 256, 0, st>>>(x, ptr, perm, N, E_cap, mean, scale, accumulate, out);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
