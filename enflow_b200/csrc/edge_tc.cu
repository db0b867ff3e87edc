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
// Thread map: 16 warps.  Accumulators are transposed (TMEM lane = hidden unit n, column = edge): warp w owns
// TMEM lanes [32 (w%4), +32) (the hardware's lane-quarter rule) and tile edges [32 (w/4), +32), i.e. in the epilogues
// every thread handles one hidden unit x 32 consecutive edges; x1 is produced one warp per edge row.
// Tiles are software-pipelined per CTA (geometry two tiles ahead, z1 gather and first MMA one tile ahead), see the
// comment in front of the tile loop and profiles/r1c_phase_times.txt.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int THREADS_V1 = 512;

struct TileInfo {
    int row[tc::TILE], col[tc::TILE], valid[tc::TILE], start[tc::TILE], mis[tc::TILE];
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
    static constexpr size_t bar_off = (t_off + 3 * sizeof(TileInfo) + 15) / 16 * 16;    // tile info: three tiles in flight
    static constexpr size_t total = bar_off + 64 + 1024;   // + alignment slack
};

// pack fp32 W [128][128] (row-major [out][in]) into swizzled bf16 images hi (and lo)
__global__ void __launch_bounds__(256) k_pack_tc(const float* __restrict__ W2, const float* __restrict__ W3,
                                                  unsigned char* __restrict__ img, int64_t param_stride,
                                                  int64_t img_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // (matrix, row, chunk16); blockIdx.y = layer
    if (idx >= 2 * 128 * 16) return;
    W2 += blockIdx.y * param_stride; W3 += blockIdx.y * param_stride; img += blockIdx.y * img_stride;
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

// 32x32 transpose-reduce across a warp: on return v[0] of lane l holds sum over lanes of (their) v[l]
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
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

// Accumulators are TRANSPOSED: TMEM lane = hidden unit n, column = edge.  Thread (n, 32 edges) therefore keeps
// its bias/weight scalars in registers, stores z rows fully coalesced (lanes = consecutive n), and writes the
// next operand as the [hidden][edge] image (16-byte vector stores) that the second GEMM reads MN-major.
template <bool SPLIT>
__global__ void __launch_bounds__(THREADS_V1, 1)
k_edge_fwd_tc_v1(const int* __restrict__ row, const int* __restrict__ col, const int* __restrict__ E_dev,
              const float* __restrict__ pos, const float* __restrict__ box, const float* __restrict__ P,
              const float* __restrict__ S, const float* __restrict__ W1, int e1, const float* __restrict__ b2,
              const float* __restrict__ b3, const float* __restrict__ wc, const unsigned char* __restrict__ wimg,
              const int* __restrict__ rowptr, const int* __restrict__ mis, float* __restrict__ runs,
              float* __restrict__ s_out, float* __restrict__ trans) {
    using L = Smem<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* Wimg = sm + L::w_off;
    unsigned char* A = sm + L::a_off;
    TileInfo* tib = reinterpret_cast<TileInfo*>(sm + L::t_off);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(sm + L::bar_off);
    uint64_t* bar_mma = bar_w + 1;
    uint64_t* bar_g1 = bar_w + 2;          // the first GEMM of a tile is issued one tile ahead: its own barrier
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int q = w & 3, cg = w >> 2;
    const int n = 32 * q + lane;           // hidden unit == TMEM lane
    const int ec = 32 * cg;                // first edge (tile-local) of this thread's 32 columns

    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_g1, 1);
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(tmem_slot, 256);
    const float b2n = b2[n], b3n = b3[n], wcn = wc[n];
    const float4 wr4 = make_float4(W1[(4 * lane + 0) * e1 + e1 - 1], W1[(4 * lane + 1) * e1 + e1 - 1],
                                   W1[(4 * lane + 2) * e1 + e1 - 1], W1[(4 * lane + 3) * e1 + e1 - 1]);
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

    const uint32_t x_base = tc::smem_u32(A);
    const uint32_t w2_base = tc::smem_u32(Wimg);
    const uint32_t w3_base = tc::smem_u32(Wimg + (size_t)(SPLIT ? 2 : 1) * tc::IMG_BYTES);
    const uint64_t dW2 = tc::make_desc(w2_base, 16, 1024), dW3 = tc::make_desc(w3_base, 16, 1024);
    const uint64_t dXk = tc::make_desc(x_base, 16, 1024), dXmn = tc::make_desc(x_base, tc::BLK_BYTES, 1024);
    const uint32_t idesc_kk = tc::make_idesc(false, false);
    const uint32_t idesc_kmn = tc::make_idesc(false, true);
    const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)ec;     // z2 accumulator; z3 is 128 columns further
    uint32_t parity = 0, parity_g1 = 0;

    const int E = E_dev[0];
    const int tiles = (E + tc::TILE - 1) / tc::TILE;
    auto geometry = [&](TileInfo& ti, int e0) {       // edge geometry (data/base.py:15-19, egcl.py:80)
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
            ti.start[tid] = ok && (e == rowptr[i]);
            ti.mis[tid] = ok ? mis[i + 1] : 0;
            ti.d[tid][0] = d0; ti.d[tid][1] = d1; ti.d[tid][2] = d2;
            ti.r[tid] = d0 * d0 + d1 * d1 + d2 * d2;
        }
    };
    // z1 = P[row] + S[col] + w_r r: one warp per edge row, lanes = 4 consecutive features (coalesced 512-byte
    // row reads); gathered one tile ahead (behind the second MMA of the previous tile) and kept in registers
    auto load_z1 = [&](const TileInfo& ti, float (&z)[tc::TILE / 16][4]) {
#pragma unroll
        for (int it = 0; it < tc::TILE / 16; ++it) {
            const int m = w + 16 * it;
            const float r = ti.r[m];
            const float4 p = __ldg(reinterpret_cast<const float4*>(P + (int64_t)ti.row[m] * ENF_H) + lane);
            const float4 s = __ldg(reinterpret_cast<const float4*>(S + (int64_t)ti.col[m] * ENF_H) + lane);
            z[it][0] = fmaf(wr4.x, r, p.x + s.x); z[it][1] = fmaf(wr4.y, r, p.y + s.y);
            z[it][2] = fmaf(wr4.z, r, p.z + s.z); z[it][3] = fmaf(wr4.w, r, p.w + s.w);
        }
    };
    // x1 = silu(z1) into the K-major [edge][feature] operand image
    auto put_x1 = [&](const float (&z)[tc::TILE / 16][4]) {
#pragma unroll
        for (int it = 0; it < tc::TILE / 16; ++it) {
            const int m = w + 16 * it;
            float x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = z[it][j] * tc::sigmoid_sfu(z[it][j]);      // padding rows: finite, masked in epilogue 1
            const uint32_t off = tc::img_chunk_offset(m, lane >> 1) + ((lane & 1) << 3);
            if (SPLIT) {
                uint2 hi, lo;
                tc::split2(x[0], x[1], hi.x, lo.x);
                tc::split2(x[2], x[3], hi.y, lo.y);
                *reinterpret_cast<uint2*>(A + off) = hi;
                *reinterpret_cast<uint2*>(A + tc::IMG_BYTES + off) = lo;
            } else {
                *reinterpret_cast<uint2*>(A + off) = make_uint2(tc::pack_bf16(x[0], x[1]), tc::pack_bf16(x[2], x[3]));
            }
        }
    };
    // Software pipeline over the CTA's tiles (t, t+1, t+2 = this CTA's consecutive tiles):
    //   geometry(t+2) and the z1 gather of t+1 run behind the second MMA of tile t;
    //   x1(t+1) is written and its first MMA issued (into the other accumulator) before epilogue 2 of tile t,
    //   so that MMA runs under the epilogue instead of in front of an idle CTA.
    auto issue_g1 = [&]() {          // z2^T[n][e] = sum_k W2[n][k] x1[e][k]: A = weight image, B = x1 image, both K-major
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            tc::issue_gemm_t<SPLIT, 8, tc::OffK128, tc::OffK128>(tmem, dW2, tc::IMG_BYTES, dXk, tc::IMG_BYTES, idesc_kk, false);
            tc::mma_commit(bar_g1);
        }
    };
    int cur = 0;
    float z1[tc::TILE / 16][4];
    const int stride = gridDim.x;
    if ((int)blockIdx.x < tiles) {
        geometry(tib[0], blockIdx.x * tc::TILE);
        if ((int)blockIdx.x + stride < tiles) geometry(tib[1], (blockIdx.x + stride) * tc::TILE);
        __syncthreads();
        load_z1(tib[0], z1);
        put_x1(z1);
        issue_g1();
    }
    for (int tile = blockIdx.x; tile < tiles; tile += stride) {
        const int e0 = tile * tc::TILE;
        TileInfo& ti = tib[cur];
        TileInfo& tn = tib[cur == 2 ? 0 : cur + 1];
        TileInfo& tnn = tib[cur == 0 ? 2 : cur - 1];
        const int next = tile + stride, next2 = next + stride;
        tc::mbar_wait(bar_g1, parity_g1);
        parity_g1 ^= 1;
        tc::fence_after_sync();
        // ---- epilogue 1: bias, x2^T = silu(z2)^T as the next operand, and the segment sums of x2 over each
        //      row (egcl.py:66) as per-run partials: this thread owns hidden unit n for 32 consecutive edges, so
        //      the sum over a row's edges is a thread-local running sum flushed at row starts (see segment.cu)
        {
            float v[32];
            tc::tmem_ld32(taddr, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] += b2n;
                v[j] = ti.valid[ec + j] ? v[j] * tc::sigmoid_sfu(v[j]) : 0.f;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int jb = 16 * half;
                if (ti.valid[ec + jb]) {
                    int rid = ((e0 + ec + jb) >> 4) + ti.mis[ec + jb];
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j > 0 && ti.start[ec + jb + j]) {
                            runs[(int64_t)rid * ENF_H + n] = acc;
                            ++rid;
                            acc = 0.f;
                        }
                        acc += v[jb + j];
                    }
                    runs[(int64_t)rid * ENF_H + n] = acc;
                }
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[8 * ch + j];
                store_chunk<SPLIT>(A, n, 4 * cg + ch, x);       // image row = hidden unit, columns = edges
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {             // z3^T = W3 x2^T  (B read MN-major)
            tc::fence_after_sync();
            // z3^T = W3 x2^T: B is the [hidden][edge] image the epilogue wrote, read MN-major
            tc::issue_gemm_t<SPLIT, 8, tc::OffK128, tc::OffMN>(tmem + 128, dW3, tc::IMG_BYTES, dXmn, tc::IMG_BYTES, idesc_kmn, false);
            tc::mma_commit(bar_mma);
        }
        if (next2 < tiles) geometry(tnn, next2 * tc::TILE);
        if (next < tiles) load_z1(tn, z1);
        tc::mbar_wait(bar_mma, parity);
        parity ^= 1;
        tc::fence_after_sync();
        if (next < tiles) {          // the operand buffer is free again: next tile's x1 and first MMA
            put_x1(z1);
            issue_g1();
        }
        // ---- epilogue 2: bias, s[e] = sum_n wc[n] silu(z3[e][n]) (transpose-reduce over the warp's 32 n)
        {
            float v[32];
            tc::tmem_ld32(taddr + 128, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] += b3n;
                v[j] = wcn * (v[j] * tc::sigmoid_sfu(v[j]));
            }
            ti.s_part[q][ec + lane] = warp_transpose_sum(v, lane);
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
        cur = cur == 2 ? 0 : cur + 1;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

int64_t enf_tc_pack_bytes() { return tc::PACK_BYTES; }

int enf_node_tc_pack(const float* lp0, int nf, int L, int64_t param_stride, unsigned char* img0, int64_t img_stride,
                     cudaStream_t st);
int enf_tc_pack_layers(const float* lp0, int nf, int L, int64_t param_stride, unsigned char* img0, int64_t img_stride,
                       cudaStream_t st);

int enf_tc_pack_layer(const float* lp, int nf, unsigned char* img, cudaStream_t st) {
    return enf_tc_pack_layers(lp, nf, 1, 0, img, 0, st);
}

// all L layers in two launches (blockIdx.y = layer): lp0 / img0 = first layer, strides in floats / bytes
int enf_tc_pack_layers(const float* lp0, int nf, int L, int64_t param_stride, unsigned char* img0, int64_t img_stride,
                       cudaStream_t st) {
    const EgclOffsets o = enf_egcl_offsets(nf);
    enf_count_launch(), k_pack_tc<<<dim3((2 * 128 * 16 + 255) / 256, L), 256, 0, st>>>(lp0 + o.off[P_W2], lp0 + o.off[P_W3], img0,
                                                                                   param_stride, img_stride);
    ENF_CHECK_LAUNCH();
    return enf_node_tc_pack(lp0, nf, L, param_stride, img0, img_stride, st);
}

int enf_edge_fwd_tc_v2(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                       const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg, int nf,
                       const int* rowptr, const int* mis, float* runs, float* s_out, float* trans, cudaStream_t st);

// mode 1 = split (fp32-accurate), mode 2 = bf16.  The product kernel is edge_tc_fwd.cu (weights in TMEM, two tiles in
// flight); ENFLOW_FWD_V1=1 selects the round-1 kernel above for A/B timing.
int enf_edge_fwd_tc(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                    const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg,
                    int nf, const int* rowptr, const int* mis, float* runs, float* s_out, float* trans,
                    cudaStream_t st) {
    if (E_cap == 0) return ENF_OK;
    static const bool v1 = getenv("ENFLOW_FWD_V1") != nullptr;
    if (!v1) return enf_edge_fwd_tc_v2(mode, row, col, E_dev, E_cap, pos, box, P, S, lp, wimg, nf, rowptr, mis, runs, s_out, trans, st);
    const EgclOffsets o = enf_egcl_offsets(nf);
    int grid = (E_cap + tc::TILE - 1) / tc::TILE;
    if (grid > enf_num_sms()) grid = enf_num_sms();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_edge_fwd_tc_v1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<true>::total);
        cudaFuncSetAttribute(k_edge_fwd_tc_v1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<false>::total);
        attr = true;
    }
    if (mode == 1)
        enf_count_launch(), k_edge_fwd_tc_v1<true><<<grid, THREADS_V1, Smem<true>::total, st>>>(
            row, col, E_dev, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3],
            lp + o.off[P_WC], wimg, rowptr, mis, runs, s_out, trans);
    else
        enf_count_launch(), k_edge_fwd_tc_v1<false><<<grid, THREADS_V1, Smem<false>::total, st>>>(
            row, col, E_dev, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, lp + o.off[P_B2], lp + o.off[P_B3],
            lp + o.off[P_WC], wimg, rowptr, mis, runs, s_out, trans);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
