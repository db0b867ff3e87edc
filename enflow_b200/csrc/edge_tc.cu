// K1 on the 5th-generation tensor cores (tcgen05 + TMEM), forward.
//
// Same math as k_edge_fwd (edge_mlp.cu; enflow/nn/egcl.py:57-63,71-75): per 128-edge tile
//   x1 = silu(P[row] + S[col] + w_r r) -> [tcgen05.mma] z2 = W2 x1 + b2 -> x2 = silu(z2)
//      -> [tcgen05.mma] z3 = W3 x2 + b3 -> s = wc . silu(z3), trans = clamp(d s, +-100)
// The activation tile is produced by the CUDA cores directly into the 128B-swizzled K-major operand image
// (tc_common.cuh), the weight images stay resident in shared memory for the life of the CTA (loaded once by
// TMA bulk copies), the accumulator lives in TMEM and is read back with tcgen05.ld for the bias + SiLU epilogue.
//
// Precision modes (profiles/r1_split_precision.txt):
//   SPLIT = true : fp32-accurate.  Operands are split x = hi + lo in bf16 and each GEMM is three MMAs
//                  (hi.hi + lo.hi + hi.lo) accumulated in fp32: end-to-end drift <= 1.3e-7 on the latents.
//   SPLIT = false: bf16 operands, one MMA per GEMM (the north star's "bf16 MLP" mode, tolerance 1e-2).
//
// Thread map: 16 warps. warp w owns TMEM lanes / edge rows [32 (w%4), +32) (the hardware's lane-quarter rule)
// and feature columns [32 (w/4), +32): every thread handles one edge row x 32 columns in all phases.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int THREADS = 512;

struct TileInfo {
    int row[tc::TILE], col[tc::TILE], valid[tc::TILE];
    float d[tc::TILE][3];
    float r[tc::TILE];
    float s_part[4][tc::TILE];
};

struct Consts {          // per-layer vectors staged once per CTA
    float wr[ENF_H], b2[ENF_H], b3[ENF_H], wc[ENF_H];
};

template <bool SPLIT>
struct Smem {
    static constexpr int NW = SPLIT ? 4 : 2;      // weight images: W2 hi[,lo], W3 hi[,lo]
    static constexpr int NA = SPLIT ? 2 : 1;      // activation images: hi[,lo]
    static constexpr size_t w_off = 0;
    static constexpr size_t a_off = (size_t)NW * tc::IMG_BYTES;
    static constexpr size_t c_off = a_off + (size_t)NA * tc::IMG_BYTES;
    static constexpr size_t t_off = c_off + sizeof(Consts);
    static constexpr size_t bar_off = (t_off + sizeof(TileInfo) + 15) / 16 * 16;
    static constexpr size_t total = bar_off + 64 + 1024;   // + alignment slack
};

// pack fp32 W [128][128] (row-major [out][in]) into swizzled bf16 images hi (and lo)
__global__ void __launch_bounds__(256) k_pack_tc(const float* __restrict__ W2, const float* __restrict__ W3,
                                                  unsigned char* __restrict__ img) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (matrix, row, chunk16)
    if (idx >= 2 * 128 * 16) return;
    const int mat = idx / (128 * 16), row = (idx / 16) % 128, ch = idx % 16;
    const float* W = (mat ? W3 : W2) + row * ENF_H + ch * 8;
    const float4 a = *reinterpret_cast<const float4*>(W), b = *reinterpret_cast<const float4*>(W + 4);
    uint4 hi, lo;
    tc::split2(a.x, a.y, hi.x, lo.x);
    tc::split2(a.z, a.w, hi.y, lo.y);
    tc::split2(b.x, b.y, hi.z, lo.z);
    tc::split2(b.z, b.w, hi.w, lo.w);
    const uint32_t off = tc::img_chunk_offset(row, ch);
    *reinterpret_cast<uint4*>(img + (size_t)(2 * mat) * tc::IMG_BYTES + off) = hi;
    *reinterpret_cast<uint4*>(img + (size_t)(2 * mat + 1) * tc::IMG_BYTES + off) = lo;
}

// write 8 consecutive columns (chunk16) of one row into the activation image(s)
template <bool SPLIT>
__device__ __forceinline__ void store_chunk(unsigned char* A, int row, int chunk16, const float (&x)[8]) {
    const uint32_t off = tc::img_chunk_offset(row, chunk16);
    if (SPLIT) {
        uint4 hi, lo;
        tc::split2(x[0], x[1], hi.x, lo.x);
        tc::split2(x[2], x[3], hi.y, lo.y);
        tc::split2(x[4], x[5], hi.z, lo.z);
        tc::split2(x[6], x[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(A + off) = hi;
        *reinterpret_cast<uint4*>(A + tc::IMG_BYTES + off) = lo;
    } else {
        uint4 hi;
        hi.x = tc::pack_bf16(x[0], x[1]); hi.y = tc::pack_bf16(x[2], x[3]);
        hi.z = tc::pack_bf16(x[4], x[5]); hi.w = tc::pack_bf16(x[6], x[7]);
        *reinterpret_cast<uint4*>(A + off) = hi;
    }
}

// D[m][n] (+)= sum_k A[m][k] W[n][k], both K-major images; SPLIT adds the two cross terms
template <bool SPLIT>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_base, uint32_t w_base, uint32_t idesc) {
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        tc::mma_f16(tmem_d, tc::desc_kmajor(a_base, ks), tc::desc_kmajor(w_base, ks), idesc, ks > 0);
    if (SPLIT) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // lo(A) . hi(W)
            tc::mma_f16(tmem_d, tc::desc_kmajor(a_base + tc::IMG_BYTES, ks), tc::desc_kmajor(w_base, ks), idesc, true);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)      // hi(A) . lo(W)
            tc::mma_f16(tmem_d, tc::desc_kmajor(a_base, ks), tc::desc_kmajor(w_base + tc::IMG_BYTES, ks), idesc, true);
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_edge_fwd_tc(const int* __restrict__ row, const int* __restrict__ col, const int* __restrict__ E_dev,
              const float* __restrict__ pos, const float* __restrict__ box, const float* __restrict__ P,
              const float* __restrict__ S, const float* __restrict__ W1, int e1, const float* __restrict__ b2,
              const float* __restrict__ b3, const float* __restrict__ wc, const unsigned char* __restrict__ wimg,
              float* __restrict__ z2, float* __restrict__ z3, float* __restrict__ s_out, float* __restrict__ trans) {
    using L = Smem<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* Wimg = sm + L::w_off;
    unsigned char* A = sm + L::a_off;
    Consts& cs = *reinterpret_cast<Consts*>(sm + L::c_off);
    TileInfo& ti = *reinterpret_cast<TileInfo*>(sm + L::t_off);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(sm + L::bar_off);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = w & 3, cg = w >> 2;
    const int m = 32 * q + lane;           // edge row in the tile == TMEM lane
    const int c0 = 32 * cg;                // first feature column of this thread

    // ---- one-time setup: barriers, TMEM, resident weight images (TMA bulk copies), per-layer vectors
    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(tmem_slot, 128);
    for (int i = tid; i < ENF_H; i += THREADS) {
        cs.wr[i] = W1[i * e1 + e1 - 1];
        cs.b2[i] = b2[i];
        cs.b3[i] = b3[i];
        cs.wc[i] = wc[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        tc::mbar_expect_tx(bar_w, L::NW * tc::IMG_BYTES);
        for (int i = 0; i < L::NW; ++i) {
            // global order: W2_hi, W2_lo, W3_hi, W3_lo ; resident order: same when SPLIT, else W2_hi, W3_hi
            const int src = SPLIT ? i : 2 * i;
            tc::bulk_g2s(Wimg + (size_t)i * tc::IMG_BYTES, wimg + (size_t)src * tc::IMG_BYTES, tc::IMG_BYTES, bar_w);
        }
    }
    tc::mbar_wait(bar_w, 0);

    const uint32_t a_base = tc::smem_u32(A);
    const uint32_t w2_base = tc::smem_u32(Wimg);
    const uint32_t w3_base = tc::smem_u32(Wimg + (size_t)(SPLIT ? 2 : 1) * tc::IMG_BYTES);
    const uint32_t idesc = tc::make_idesc(false, false);
    const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0;
    uint32_t parity = 0;

    const int E = E_dev[0];
    const int tiles = (E + tc::TILE - 1) / tc::TILE;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int e0 = tile * tc::TILE;
        // ---- edge geometry (data/base.py:15-19, egcl.py:80)
        if (tid < tc::TILE) {
            const int e = e0 + tid;
            const bool ok = e < E;
            int i = 0, j = 0;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (ok) {
                i = row[e]; j = col[e];
                d0 = wrapf_(pos[(int64_t)i * 3 + 0] - pos[(int64_t)j * 3 + 0], 0.5f * box[(int64_t)i * 3 + 0]);
                d1 = wrapf_(pos[(int64_t)i * 3 + 1] - pos[(int64_t)j * 3 + 1], 0.5f * box[(int64_t)i * 3 + 1]);
                d2 = wrapf_(pos[(int64_t)i * 3 + 2] - pos[(int64_t)j * 3 + 2], 0.5f * box[(int64_t)i * 3 + 2]);
            }
            ti.row[tid] = i; ti.col[tid] = j; ti.valid[tid] = ok;
            ti.d[tid][0] = d0; ti.d[tid][1] = d1; ti.d[tid][2] = d2;
            ti.r[tid] = d0 * d0 + d1 * d1 + d2 * d2;
        }
        __syncthreads();
        const bool ok = ti.valid[m];
        // ---- x1 = silu(P[row] + S[col] + w_r r) into the operand image
        {
            const float r = ti.r[m];
            const float* p = P + (int64_t)ti.row[m] * ENF_H + c0;
            const float* s = S + (int64_t)ti.col[m] * ENF_H + c0;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(p + 8 * ch));
                const float4 p1 = __ldg(reinterpret_cast<const float4*>(p + 8 * ch + 4));
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + 8 * ch));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + 8 * ch + 4));
                const float4 w0 = *reinterpret_cast<const float4*>(cs.wr + c0 + 8 * ch);
                const float4 w1 = *reinterpret_cast<const float4*>(cs.wr + c0 + 8 * ch + 4);
                float x[8] = {fmaf(w0.x, r, p0.x + s0.x), fmaf(w0.y, r, p0.y + s0.y), fmaf(w0.z, r, p0.z + s0.z),
                              fmaf(w0.w, r, p0.w + s0.w), fmaf(w1.x, r, p1.x + s1.x), fmaf(w1.y, r, p1.y + s1.y),
                              fmaf(w1.z, r, p1.z + s1.z), fmaf(w1.w, r, p1.w + s1.w)};
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = ok ? x[j] * tc::sigmoid_sfu(x[j]) : 0.f;
                store_chunk<SPLIT>(A, m, 4 * cg + ch, x);
            }
        }
        tc::fence_async_smem();
        __syncthreads();
        // ---- z2 = x1 W2^T on the tensor core
        if (tid == 0) {
            tc::fence_after_sync();
            issue_gemm<SPLIT>(tmem, a_base, w2_base, idesc);
            tc::mma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
        // ---- epilogue 1: bias, save z2, x2 = silu(z2) back into the operand image
        {
            float v[32];
            tc::tmem_ld32(taddr, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += cs.b2[c0 + j];
            if (ok) {
                float4* dst = reinterpret_cast<float4*>(z2 + (int64_t)(e0 + m) * ENF_H + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float z = v[8 * ch + j];
                    x[j] = ok ? z * tc::sigmoid_sfu(z) : 0.f;
                }
                store_chunk<SPLIT>(A, m, 4 * cg + ch, x);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        // ---- z3 = x2 W3^T
        if (tid == 0) {
            tc::fence_after_sync();
            issue_gemm<SPLIT>(tmem, a_base, w3_base, idesc);
            tc::mma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
        // ---- epilogue 2: bias, save z3, s = wc . silu(z3)
        {
            float v[32];
            tc::tmem_ld32(taddr, v);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] += cs.b3[c0 + j];
                part = fmaf(cs.wc[c0 + j], v[j] * tc::sigmoid_sfu(v[j]), part);
            }
            if (ok) {
                float4* dst = reinterpret_cast<float4*>(z3 + (int64_t)(e0 + m) * ENF_H + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            ti.s_part[cg][m] = part;
        }
        tc::fence_before_sync();
        __syncthreads();
        if (tid < tc::TILE && ti.valid[tid]) {
            const int e = e0 + tid;
            const float s = (ti.s_part[0][tid] + ti.s_part[1][tid]) + (ti.s_part[2][tid] + ti.s_part[3][tid]);
            s_out[e] = s;
#pragma unroll
            for (int c = 0; c < 3; ++c) trans[(int64_t)e * 3 + c] = fminf(fmaxf(ti.d[tid][c] * s, -100.f), 100.f);
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, 128);
}

}  // namespace

int64_t enf_tc_pack_bytes() { return 4 * (int64_t)tc::IMG_BYTES; }

int enf_tc_pack_layer(const float* lp, int nf, unsigned char* img, cudaStream_t st) {
    const EgclOffsets o = enf_egcl_offsets(nf);
    enf_count_launch(), k_pack_tc<<<(2 * 128 * 16 + 255) / 256, 256, 0, st>>>(lp + o.off[P_W2], lp + o.off[P_W3], img);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}

// mode 1 = split (fp32-accurate), mode 2 = bf16
int enf_edge_fwd_tc(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                    const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg,
                    int nf, float* z2, float* z3, float* s_out, float* trans, cudaStream_t st) {
    if (E_cap == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    int grid = (E_cap + tc::TILE - 1) / tc::TILE;
    if (grid > enf_num_sms()) grid = enf_num_sms();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_edge_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<true>::total);
        cudaFuncSetAttribute(k_edge_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<false>::total);
        attr = true;
    }
    if (mode == 1)
        enf_count_launch(), k_edge_fwd_tc<true><<<grid, THREADS, Smem<true>::total, st>>>(
            row, col, E_dev, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3],
            lp + o.off[P_WC], wimg, z2, z3, s_out, trans);
    else
        enf_count_launch(), k_edge_fwd_tc<false><<<grid, THREADS, Smem<false>::total, st>>>(
            row, col, E_dev, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3],
            lp + o.off[P_WC], wimg, z2, z3, s_out, trans);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
