// K1 forward on tcgen05, round-2 structure: WEIGHTS IN TENSOR MEMORY, two tiles in flight, warp-specialised.
//
// Math (enflow/nn/egcl.py:57-63,71-75), per 128-edge tile:
//   x1 = silu(P[row] + S[col] + w_r r) -> G1: z2^T = W2 x1^T -> x2 = silu(z2 + b2), row sums of x2 as run partials
//      -> G2: z3^T = W3 x2^T -> s = wc . silu(z3 + b3), trans = clamp(d s, +-100)
//
// Why this shape (profiles/r2a_edge_fwd_tc_*.txt, profiles/README.md): the round-1 kernel kept the four weight images
// (128 KB, fp32-accurate mode) in shared memory, which left room for ONE activation buffer, so a tile's second GEMM
// ran in front of an idle CTA and nothing was above 50 % busy (issue 50 %, XU 46 %, L1 data pipe 57 %, tensor 23 %).
// Here the weights are the A operand read from TENSOR MEMORY (tcgen05.mma with A in TMEM): thread (hidden unit n)
// moves row n of the packed bf16 images of W2 / W3 (hi and lo, 64 columns each) into TMEM with tcgen05.st once per CTA.  That
//   * frees 128 KB of shared memory: two 64 KB activation buffers, i.e. two 128-edge tiles in flight at N = 128;
//   * halves the MMA operand traffic on the L1 data pipe (only the 4 KB activation slice is read per MMA);
//   * keeps the transposed accumulator (TMEM lane = hidden unit, column = edge) the reductions over edges need.
// TMEM (512 columns): T_A [0,128) T_B [128,256) | W2 hi, W2 lo, W3 hi, W3 lo [256,512) (bf16 mode: W2, W3).
//
// Warps: 0..15 epilogue (warp w: TMEM lanes 32 (w % 4) + lane = hidden unit n, tile edges [32 (w / 4), +32)),
// 16 issues every tcgen05.mma, 17..19 produce the per-tile edge records (indices, wrapped differences, |d|^2, run
// headers) into a 4-slot shared-memory ring, up to four tiles ahead.  No block barrier inside the tile loop: tasks end
// with one mbarrier arrival per warp, the issuing warp waits for all 16 and issues the next GEMM of that pipeline.
// A CTA's tiles alternate between pipelines A (even) and B (odd); every epilogue warp runs
//      E1(a) | E2(b - 2) + X(b) | E2(a) + X(a + 2) | E1(b) | ...        (X = gather + x1 image)
// so the GEMM a task waits for was issued one task earlier and ran under the other pipeline's epilogue.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int EPI_THREADS = 512;
constexpr int THREADS = EPI_THREADS + 128;
// 20 warps = 5 per scheduler partition -> 96 registers per thread at launch; the control warpgroup gives registers back
// (setmaxnreg.dec) and the epilogue warpgroups grow: 512 * 104 + 128 * 56 <= 640 * 96
constexpr int EPI_REGS = 104, CTRL_REGS = 56;
constexpr int TE = tc::TILE;                 // 128 edges per tile
constexpr int RING = 4;
constexpr uint32_t W_COL = 256;              // first TMEM column of the weight images

struct TileRec {                             // written by the geometry warps, read by the epilogue warps
    int row[TE], col[TE];
    float r[TE];
    float d[TE][3];
    int2 hdr[TE / 16];                       // per 16-edge group: run id of its first edge (-1: none valid), row-start bits
};

template <bool SPLIT>
struct SmemF {
    static constexpr int NA = SPLIT ? 2 : 1;                       // images per activation buffer (hi[, lo])
    static constexpr size_t XBUF = (size_t)NA * tc::IMG_BYTES;
    static constexpr size_t x_off = 0;
    static constexpr size_t rec_off = 2 * XBUF;
    static constexpr size_t sp_off = rec_off + RING * sizeof(TileRec);
    static constexpr size_t bar_off = (sp_off + 2 * 4 * TE * sizeof(float) + 15) / 16 * 16;
    static constexpr size_t total = bar_off + 256 + 1024;          // + alignment slack
};

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// One GEMM with A in TMEM: 1 (bf16) or 3 (hi.hi, hi.lo, lo.hi) passes over 2 blocks of four k-steps.  A advances 8 columns
// per k-step (16 bf16 = 8 words per lane); k-step (blk, i) of B is at blk * B_BLK + i * B_IN bytes.
template <bool SPLIT, uint32_t B_BLK, uint32_t B_IN>
__device__ __forceinline__ void issue_gemm_ts(bool leader, uint32_t tmem_d, uint32_t a, uint64_t b, uint32_t b_lo_bytes,
                                              uint32_t idesc) {
    constexpr int NP = SPLIT ? 3 : 1;
#pragma unroll 1
    for (int p = 0; p < NP; ++p) {
        uint32_t ap = a + (p == 2 ? 64u : 0u);
        uint64_t bp = b + (uint64_t)((p == 1 ? b_lo_bytes : 0u) >> 4);
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool acc = p > 0 || blk > 0 || i > 0;
                if (leader) mma_f16_ts(tmem_d, ap + 8u * i, bp + (uint64_t)((i * B_IN) >> 4), idesc, acc);
            }
            ap += 32u;
            bp += (uint64_t)(B_BLK >> 4);
        }
    }
}

// 8 consecutive edge columns of image row n ([hidden][edge] image, 128 rows x 128 columns)
template <bool SPLIT>
__device__ __forceinline__ void store_chunk(unsigned char* A, int n, int chunk16, const float (&x)[8]) {
    const uint32_t off = tc::img_chunk_offset(n, chunk16);
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

template <int P> struct Pipe { static constexpr int value = P; };

template <bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1)
k_edge_fwd_tc(const int* __restrict__ row, const int* __restrict__ col, const int* __restrict__ E_dev, int E_cap,
              const float* __restrict__ pos, const float* __restrict__ box, const float* __restrict__ P,
              const float* __restrict__ S, const float* __restrict__ W1, int e1,
              const unsigned char* __restrict__ wimg, const float* __restrict__ b2, const float* __restrict__ b3,
              const float* __restrict__ wc, const int* __restrict__ rowptr, const int* __restrict__ mis,
              float* __restrict__ runs, float* __restrict__ s_out, float* __restrict__ trans) {
    using L = SmemF<SPLIT>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024 - (tc::smem_u32(smem_raw) & 1023)) & 1023);
    unsigned char* XB = sm + L::x_off;                  // XB + p XBUF: x1 image [edge][k], then x2^T image [k][edge]
    TileRec* rec = reinterpret_cast<TileRec*>(sm + L::rec_off);
    float* s_part = reinterpret_cast<float*>(sm + L::sp_off);          // [pipeline][quarter][edge]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::bar_off);
    uint64_t* bar_acc = bars;                  // [2] accumulator of pipeline p is ready             (MMA -> epilogue)
    uint64_t* bar_opnd = bars + 2;             // [2] operand image of pipeline p is written, T read (epilogue -> MMA)
    uint64_t* bar_full = bars + 4;             // [RING] tile record written                          (geometry -> epilogue)
    uint64_t* bar_empty = bars + 4 + RING;     // [RING] tile record no longer read                   (epilogue -> geometry)
    uint64_t* bar_w = bars + 4 + 2 * RING;      // weight images have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * RING);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        for (int p = 0; p < 2; ++p) {
            tc::mbar_init(bar_acc + p, 1);
            tc::mbar_init(bar_opnd + p, EPI_THREADS / 32);
        }
        for (int s = 0; s < RING; ++s) {
            tc::mbar_init(bar_full + s, TE / 32);              // one arrival per 32-edge group
            tc::mbar_init(bar_empty + s, EPI_THREADS / 32);
        }
        tc::mbar_fence_init();
    }
    __syncwarp();
    if (w == 0) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int E = E_dev[0] < E_cap ? E_dev[0] : E_cap;
    const int tiles = (E + TE - 1) / TE;
    // this CTA's tiles: tile0 .. tile0 + T - 1 (a contiguous range keeps a molecule's rows of P / S in this SM's L1)
    const int Tc = (tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile0 = blockIdx.x * Tc;
    const int T = tiles - tile0 < 0 ? 0 : (tiles - tile0 < Tc ? tiles - tile0 : Tc);

    // ---- weights -> TMEM (A operand).  The packed bf16 images (k_pack_tc: W2 hi, W2 lo, W3 hi, W3 lo, each [n][k] in the
    //      swizzled operand layout) are staged through the still unused activation buffers by TMA bulk copies; thread n
    //      then moves row n of one image (256 B = 64 packed words) into 64 TMEM columns with tcgen05.st.
    constexpr int NW = SPLIT ? 4 : 2;
    if (tid == 0) {
        tc::mbar_expect_tx(bar_w, NW * tc::IMG_BYTES);
        for (int i = 0; i < NW; ++i)
            tc::bulk_g2s(XB + (size_t)i * tc::IMG_BYTES, wimg + (size_t)(SPLIT ? i : 2 * i) * tc::IMG_BYTES, tc::IMG_BYTES, bar_w);
    }
    if (w < 4 * NW) {
        const int q = w & 3, img = w >> 2;
        const int n = 32 * q + lane;
        tc::mbar_wait(bar_w, 0);
        const unsigned char* src = XB + (size_t)img * tc::IMG_BYTES;
        const uint32_t dst = tmem + ((uint32_t)(32 * q) << 16) + W_COL + 64u * img;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                      // 32 input features = 4 chunks of 16 bytes = 16 packed words
            float v[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 x = *reinterpret_cast<const uint4*>(src + tc::img_chunk_offset(n, 4 * c + j));
                v[4 * j] = __uint_as_float(x.x); v[4 * j + 1] = __uint_as_float(x.y);
                v[4 * j + 2] = __uint_as_float(x.z); v[4 * j + 3] = __uint_as_float(x.w);
            }
            tc::tmem_st16(dst + 16u * c, v);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    if (w >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
        if (w == 16) {
            // =========================== MMA issuing warp ===========================
            const uint32_t idesc_kk = tc::make_idesc(false, false);       // G1: B = x1 image, K-major
            const uint32_t idesc_kmn = tc::make_idesc(false, true);       // G2: B = x2^T image, MN-major
            const uint32_t w2 = tmem + W_COL, w3 = tmem + W_COL + (SPLIT ? 128u : 64u);
            uint32_t ph[2] = {0, 0};
            auto step = [&](int p, bool second) {
                tc::mbar_wait(bar_opnd + p, ph[p] & 1);
                ++ph[p];
                tc::fence_after_sync();
                const bool leader = tc::elect_one();
                const uint32_t xb = tc::smem_u32(XB + p * L::XBUF);
                if (!second)
                    issue_gemm_ts<SPLIT, tc::BLK_BYTES, 32>(leader, tmem + 128u * p, w2, tc::make_desc(xb, 16, 1024),
                                                            tc::IMG_BYTES, idesc_kk);
                else
                    issue_gemm_ts<SPLIT, 8192, 2048>(leader, tmem + 128u * p, w3, tc::make_desc(xb, tc::BLK_BYTES, 1024),
                                                     tc::IMG_BYTES, idesc_kmn);
                if (leader) tc::mma_commit(bar_acc + p);
                __syncwarp();
            };
            if (T > 0) step(0, false);                                     // X(0) -> G1(0)
            const int periods = T / 2 + 1;
            for (int k = 0; k < periods; ++k) {
                const int a = 2 * k, b = a + 1;
                if (a < T) step(0, true);                                  // E1(a)           -> G2(a)
                if (b < T) step(1, false);                                 // E2(b-2) + X(b)  -> G1(b)
                if (a + 2 < T) step(0, false);                             // E2(a) + X(a+2)  -> G1(a+2)
                if (b < T) step(1, true);                                  // E1(b)           -> G2(b)
            }
        } else {
            // =========================== geometry warps (data/base.py:15-19, egcl.py:80) ===========================
            for (int gi = w - 17; gi < 4 * T; gi += 3) {
                const int t = gi >> 2, g = gi & 3, slot = t % RING;
                if (t >= RING) tc::mbar_wait(bar_empty + slot, (uint32_t)(t / RING - 1) & 1);
                TileRec& tr = rec[slot];
                const int m = 32 * g + lane;
                const int e = (tile0 + t) * TE + m;
                const bool ok = e < E;
                int i = 0, j = 0, start = 0, mi = 0;
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;
                if (ok) {
                    i = row[e]; j = col[e];
                    d0 = wrapf_(pos[(int64_t)i * 3 + 0] - pos[(int64_t)j * 3 + 0], 0.5f * box[(int64_t)i * 3 + 0]);
                    d1 = wrapf_(pos[(int64_t)i * 3 + 1] - pos[(int64_t)j * 3 + 1], 0.5f * box[(int64_t)i * 3 + 1]);
                    d2 = wrapf_(pos[(int64_t)i * 3 + 2] - pos[(int64_t)j * 3 + 2], 0.5f * box[(int64_t)i * 3 + 2]);
                    start = e == rowptr[i];
                    mi = mis[i + 1];
                }
                tr.row[m] = i; tr.col[m] = j;
                tr.r[m] = d0 * d0 + d1 * d1 + d2 * d2;
                tr.d[m][0] = d0; tr.d[m][1] = d1; tr.d[m][2] = d2;
                const unsigned bits = __ballot_sync(0xffffffffu, start);
                if ((lane & 15) == 0) tr.hdr[2 * g + (lane >> 4)] = make_int2(ok ? (e >> 4) + mi : -1, (int)((bits >> (lane & 16)) & 0xffffu));
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar_full + slot);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
        // =========================== epilogue warps ===========================
        const int q = w & 3, cg = w >> 2;
        const int n = 32 * q + lane;           // hidden unit == TMEM lane
        const int ec = 32 * cg;                // first tile edge of this thread's 32 columns
        const float b2n = b2[n], b3n = b3[n], wcn = wc[n];
        const float4 wr4 = make_float4(W1[(4 * lane + 0) * e1 + e1 - 1], W1[(4 * lane + 1) * e1 + e1 - 1],
                                       W1[(4 * lane + 2) * e1 + e1 - 1], W1[(4 * lane + 3) * e1 + e1 - 1]);
        const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);

        auto wait_acc = [&](int p, uint32_t parity) {
            tc::mbar_wait(bar_acc + p, parity);
            tc::fence_after_sync();
        };
        auto arrive = [&](uint64_t* bar) {     // this warp's shared-memory writes and TMEM reads of the task are done
            tc::fence_before_sync();
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar);
        };
        // z1 = P[row] + S[col] + w_r r for the warp's 8 consecutive tile edges, lanes = 4 consecutive features (512-byte
        // row reads).  CSR order: the 8 edges almost always span at most two rows, so P is requested for the first and the
        // last edge only and re-requested for an edge in between only if its row is neither.
        // the rows the gather will read are pulled into L1 one task phase ahead (no destination registers: 96 per thread do not
        // hold the gathered rows next to the accumulator values of the epilogue in between)
        auto prefetch_z1 = [&](int t) {
            const int slot = t % RING;
            tc::mbar_wait(bar_full + slot, (uint32_t)(t / RING) & 1);
            const TileRec& tr = rec[slot];
            const int4 ca = *reinterpret_cast<const int4*>(tr.col + 8 * w), cb = *reinterpret_cast<const int4*>(tr.col + 8 * w + 4);
            const int cl[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int it = 0; it < 8; ++it)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float4*>(S + (int64_t)cl[it] * ENF_H) + lane));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float4*>(P + (int64_t)tr.row[8 * w] * ENF_H) + lane));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float4*>(P + (int64_t)tr.row[8 * w + 7] * ENF_H) + lane));
        };
        auto gather_z1 = [&](int t, float4 (&z)[8]) {
            const TileRec& tr = rec[t % RING];
            {
                const int4 ca = *reinterpret_cast<const int4*>(tr.col + 8 * w), cb = *reinterpret_cast<const int4*>(tr.col + 8 * w + 4);
                const int cl[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
                for (int it = 0; it < 8; ++it) z[it] = __ldg(reinterpret_cast<const float4*>(S + (int64_t)cl[it] * ENF_H) + lane);
            }
            const int4 ra = *reinterpret_cast<const int4*>(tr.row + 8 * w), rb = *reinterpret_cast<const int4*>(tr.row + 8 * w + 4);
            const int rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
            const float4 pf = __ldg(reinterpret_cast<const float4*>(P + (int64_t)rw[0] * ENF_H) + lane);
            const float4 pl = __ldg(reinterpret_cast<const float4*>(P + (int64_t)rw[7] * ENF_H) + lane);
            const float4 r0 = *reinterpret_cast<const float4*>(tr.r + 8 * w), r1 = *reinterpret_cast<const float4*>(tr.r + 8 * w + 4);
            const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                float4 p = rw[it] == rw[0] ? pf : pl;
                if (rw[it] != rw[0] && rw[it] != rw[7]) p = __ldg(reinterpret_cast<const float4*>(P + (int64_t)rw[it] * ENF_H) + lane);
                z[it].x = fmaf(wr4.x, rr[it], p.x + z[it].x); z[it].y = fmaf(wr4.y, rr[it], p.y + z[it].y);
                z[it].z = fmaf(wr4.z, rr[it], p.z + z[it].z); z[it].w = fmaf(wr4.w, rr[it], p.w + z[it].w);
            }
        };
        // x1 = silu(z1) into the K-major [edge][feature] image of pipeline p (padding rows: finite, never summed)
        auto put_x1 = [&](unsigned char* X, const float4 (&z)[8]) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int m = 8 * w + it;
                const float x0 = z[it].x * tc::sigmoid_t<SPLIT>(z[it].x), x1 = z[it].y * tc::sigmoid_t<SPLIT>(z[it].y);
                const float x2 = z[it].z * tc::sigmoid_t<SPLIT>(z[it].z), x3 = z[it].w * tc::sigmoid_t<SPLIT>(z[it].w);
                const uint32_t off = tc::img_chunk_offset(m, lane >> 1) + ((lane & 1) << 3);
                if (SPLIT) {
                    uint2 hi, lo;
                    tc::split2(x0, x1, hi.x, lo.x);
                    tc::split2(x2, x3, hi.y, lo.y);
                    *reinterpret_cast<uint2*>(X + off) = hi;
                    *reinterpret_cast<uint2*>(X + tc::IMG_BYTES + off) = lo;
                } else {
                    *reinterpret_cast<uint2*>(X + off) = make_uint2(tc::pack_bf16(x0, x1), tc::pack_bf16(x2, x3));
                }
            }
        };
        // ---- E1(t): x2^T = silu(z2 + b2)^T over the x1 image, and the row sums of x2 (egcl.py:66) as per-run partials:
        //      the thread owns hidden unit n for 32 consecutive edges, so the sum over a row's edges is a thread-local
        //      running sum flushed at row starts (segment.cu)
        auto taskE1 = [&](auto PP, int t) {
            constexpr int p = decltype(PP)::value;
            const TileRec& tr = rec[t % RING];
            const int e0 = (tile0 + t) * TE;
            const int2 hd[2] = {tr.hdr[2 * cg], tr.hdr[2 * cg + 1]};
            wait_acc(p, 0);
            float v[32];
            tc::tmem_ld32(lane_base + 128u * p + ec, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float z = v[j] + b2n;
                v[j] = z * tc::sigmoid_t<SPLIT>(z);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float x[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) x[j] = v[16 * half + j];
                tc::run_sums16(x, hd[half].x, (unsigned)hd[half].y, E - (e0 + ec + 16 * half), runs, n);
            }
            unsigned char* X = XB + p * L::XBUF;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v[8 * ch + j];
                store_chunk<SPLIT>(X, n, 4 * cg + ch, x);
            }
            arrive(bar_opnd + p);
        };
        // ---- E2(t) + X(t2): s = wc . silu(z3 + b3) (32 x 16 transpose-reduce per warp, the four quarters meet in shared
        //      memory), trans = clamp(d s); in between, the x1 image of this pipeline's next tile, so that its first GEMM is
        //      issued before the output stores
        auto taskE2X = [&](auto PP, int t, int t2) {
            constexpr int p = decltype(PP)::value;
            const bool has_t = t >= 0 && t < T, has_x = t2 < T;
            if (has_x) prefetch_z1(t2);
            float* sp = s_part + p * 4 * TE;
            if (has_t) {
                wait_acc(p, 1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[16];
                    tc::tmem_ld16(lane_base + 128u * p + ec + 16 * half, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float y = v[j] + b3n;
                        v[j] = wcn * (y * tc::sigmoid_t<SPLIT>(y));
                    }
                    const float tsum = tc::warp_transpose_sum16(v, lane);
                    if (!(lane & 1)) sp[q * TE + ec + 16 * half + (lane >> 1)] = tsum;
                }
            }
            if (has_x) {                                           // G2(t) has completed: the buffer is free
                float4 z[8];
                gather_z1(t2, z);
                put_x1(XB + p * L::XBUF, z);
                arrive(bar_opnd + p);                              // (also orders this warp's TMEM reads in front of G1(t2))
            }
            if (has_t) {
                const int slot = t % RING;
                const TileRec& tr = rec[slot];
                const int e0 = (tile0 + t) * TE;
                tc::named_bar_sync(1 + cg, 128);                   // the four quarter-warps of this edge group
                if (lane < 24) {
                    const int el = ec + 8 * q + lane / 3, c = lane % 3;
                    const int e = e0 + el;
                    if (e < E) {
                        const float s = (sp[el] + sp[TE + el]) + (sp[2 * TE + el] + sp[3 * TE + el]);
                        trans[(int64_t)e * 3 + c] = fminf(fmaxf(tr.d[el][c] * s, -100.f), 100.f);      // egcl.py:73
                        if (c == 0) s_out[e] = s;
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bar_empty + slot);  // last reader of this tile's record
            }
        };

        taskE2X(Pipe<0>{}, -1, 0);                                 // X(0)
        const int periods = T / 2 + 1;
        for (int k = 0; k < periods; ++k) {
            const int a = 2 * k, b = a + 1;
            if (a < T) taskE1(Pipe<0>{}, a);
            taskE2X(Pipe<1>{}, b - 2, b);
            taskE2X(Pipe<0>{}, a, a + 2);
            if (b < T) taskE1(Pipe<1>{}, b);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (w == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

int enf_edge_fwd_tc_v2(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                       const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg, int nf,
                       const int* rowptr, const int* mis, float* runs, float* s_out, float* trans, cudaStream_t st) {
    if (E_cap == 0) return ENF_OK;
    const EgclOffsets o = enf_egcl_offsets(nf);
    int grid = (E_cap + TE - 1) / TE;
    if (grid > enf_num_sms()) grid = enf_num_sms();
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_edge_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemF<true>::total);
        cudaFuncSetAttribute(k_edge_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemF<false>::total);
        attr = true;
    }
    if (mode == 1)
        enf_count_launch(), k_edge_fwd_tc<true><<<grid, THREADS, SmemF<true>::total, st>>>(
            row, col, E_dev, E_cap, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, wimg,
            lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], rowptr, mis, runs, s_out, trans);
    else
        enf_count_launch(), k_edge_fwd_tc<false><<<grid, THREADS, SmemF<false>::total, st>>>(
            row, col, E_dev, E_cap, pos, box, P, S, lp + o.off[P_W1], 2 * nf + 1, wimg,
            lp + o.off[P_B2], lp + o.off[P_B3], lp + o.off[P_WC], rowptr, mis, runs, s_out, trans);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
