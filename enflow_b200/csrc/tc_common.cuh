// tcgen05 / TMEM / mbarrier / bulk-copy primitives (inline PTX, sm_100a) shared by the tensor-core kernels.
//
// Operand images in shared memory: a [128 rows][128 cols] bf16 tile is stored as two 16 KB blocks of 64
// columns; inside a block each row is 128 B and the 16-byte chunk index is XOR-ed with (row % 8)
// (the hardware SWIZZLE_128B pattern, blocks 1024-byte aligned).  Read as a K-major operand the rows are the
// M/N index and the columns are K; read as an MN-major operand the rows are K and the columns M/N.  The same
// bytes serve both views, which is what lets one weight image feed the forward GEMM and its dgrad.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

constexpr int TILE = 128;
constexpr int BLK_BYTES = 128 * 128;          // one 64-column block of a 128-row tile
constexpr int IMG_BYTES = 2 * BLK_BYTES;      // 128 x 128 bf16

// Per-layer weight image buffer (bytes): the edge kernels' W2 hi, W2 lo, W3 hi, W3 lo, then the node kernels'
// images (node_tc.cu), each as hi followed by lo.
constexpr int NODE_W4A = 4 * IMG_BYTES;                  // node_nn.0.weight[:, nf:]  [k][jj]      2 x 32 KB
constexpr int NODE_W4H = NODE_W4A + 2 * IMG_BYTES;       // node_nn.0.weight[:, :nf]^T [c 16][k]   2 x 4 KB
constexpr int NODE_W4F = NODE_W4H + 2 * 4096;            // node_nn.0.weight [k][agg | h | pad 192] 2 x 48 KB
constexpr int NODE_W5 = NODE_W4F + 2 * 3 * BLK_BYTES;    // node_nn.2.weight [c 16][k]             2 x 4 KB
constexpr int PACK_BYTES = NODE_W5 + 2 * 4096;

// byte offset of the 16-byte chunk holding columns [8*chunk16, 8*chunk16+8) of `row`
__host__ __device__ inline uint32_t img_chunk_offset(int row, int chunk16) {
    const int blk = chunk16 >> 3, c = chunk16 & 7;
    return (uint32_t)(blk * BLK_BYTES + row * 128 + ((c ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout) --------------------------
//  [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//  [46,48) version = 1 (sm_100)   [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major view, k-step ks (16 columns each) of a tile image at `base`: 32 bytes per step inside a 64-column
// block, next block 16 KB further; 8-row groups are 1024 B apart.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int ks) {
    return make_desc(base + (ks >> 2) * BLK_BYTES + (ks & 3) * 32, 16, 1024);
}
// MN-major view (rows = K): k-step ks covers rows [16 ks, 16 ks + 16) = two 8-row groups (SBO = 1024 B);
// the two 64-wide M/N blocks are LBO = 16 KB apart.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int ks) {
    return make_desc(base + ks * 2048, BLK_BYTES, 1024);
}

// general forms: `blk_stride` = bytes between the 64-column blocks of the image (rows * 128)
__device__ __forceinline__ uint64_t desc_k(uint32_t base, int ks, uint32_t blk_stride) {
    return make_desc(base + (ks >> 2) * blk_stride + (ks & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int ks, uint32_t blk_stride) {
    return make_desc(base + ks * 2048, blk_stride, 1024);
}

// ---- instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> fp32, M = N = 128 --------------
__device__ __forceinline__ uint32_t make_idesc(bool a_mn_major, bool b_mn_major, uint32_t n = 128) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format = F32
    d |= 1u << 7;                       // a_format = BF16
    d |= 1u << 10;                      // b_format = BF16
    d |= (a_mn_major ? 1u : 0u) << 15;
    d |= (b_mn_major ? 1u : 0u) << 16;
    d |= (n >> 3) << 17;                // N
    d |= (128u >> 4) << 24;             // M
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cheap MMA issue: base descriptors are built once, per-k-step offsets are compile-time constants ---------
// (the start-address field counts 16-byte units and never carries out of its 14 bits for a 227 KB window)
struct OffK128 { static constexpr uint32_t off(int ks) { return (uint32_t)((ks >> 2) * BLK_BYTES + (ks & 3) * 32); } };  // K-major, 128-row image
struct OffK64  { static constexpr uint32_t off(int ks) { return (uint32_t)((ks >> 2) * 8192 + (ks & 3) * 32); } };        // K-major, 64-row image
struct OffK    { static constexpr uint32_t off(int ks) { return (uint32_t)(ks * 32); } };                                   // K-major inside one 64-col block
struct OffMN   { static constexpr uint32_t off(int ks) { return (uint32_t)(ks * 2048); } };                                 // MN-major: 16 rows per k-step
__device__ __forceinline__ uint64_t desc_at(uint64_t d, uint32_t off_bytes) { return d + (uint64_t)(off_bytes >> 4); }

// one GEMM = 1 (bf16) or 3 (bf16x3 split: hi.hi, hi.lo, lo.hi) chains of KS MMAs, fully unrolled
template <bool SPLIT, int KS, typename OA, typename OB>
__device__ __forceinline__ void issue_gemm_t(uint32_t tmem_d, uint64_t a, uint32_t a_lo_bytes, uint64_t b,
                                             uint32_t b_lo_bytes, uint32_t idesc, bool acc_first);

// exactly one lane of the (converged) warp gets true
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// Compact issue loop for a dedicated MMA-issuing warp: the whole warp runs the (uniform) loop and address arithmetic,
// only the tcgen05.mma itself is predicated on `leader` (from elect_one).  One GEMM = 1 (bf16) or 3 (bf16x3 split)
// passes over NBLK blocks of four k-steps; k-step (blk, i) of an operand is at blk * BLK + i * IN bytes.
// Small code (a few dozen instructions per call site) matters: the fully unrolled form of a 10-GEMM tile loop is
// ~100 KB of straight-line code that is executed once per pass and misses the instruction cache on every line.
template <bool SPLIT, int NBLK, uint32_t A_BLK, uint32_t A_IN, uint32_t B_BLK, uint32_t B_IN>
__device__ __forceinline__ void issue_gemm_loop(bool leader, uint32_t tmem_d, uint64_t a, uint32_t a_lo_bytes, uint64_t b,
                                                uint32_t b_lo_bytes, uint32_t idesc, bool acc_first) {
    constexpr int NP = SPLIT ? 3 : 1;
#pragma unroll 1
    for (int p = 0; p < NP; ++p) {
        uint64_t ap = a + (uint64_t)((p == 2 ? a_lo_bytes : 0u) >> 4);
        uint64_t bp = b + (uint64_t)((p == 1 ? b_lo_bytes : 0u) >> 4);
#pragma unroll 1
        for (int blk = 0; blk < NBLK; ++blk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool acc = acc_first || p > 0 || blk > 0 || i > 0;
                if (leader) mma_f16(tmem_d, ap + (uint64_t)((i * A_IN) >> 4), bp + (uint64_t)((i * B_IN) >> 4), idesc, acc);
            }
            ap += (uint64_t)(A_BLK >> 4);
            bp += (uint64_t)(B_BLK >> 4);
        }
    }
}

// ---- TMEM ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {         // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: lane (quarter base + laneid), v[j] = column (col0 + j)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// registers -> TMEM, same shape as tmem_ld16 (per-thread scratch in spare accumulator columns); completes before return
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// named barrier over `nthreads` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- bf16 splitting ----------------------------------------------------------------------------------------
// x = hi + lo (+ O(2^-17 |x|)): hi = bf16(x), lo = bf16(x - hi).  Packs two values per 32-bit word.
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// sigmoid through the SFU: 1 / (1 + 2^(-x log2 e)); ~2 ulp, two MUFU ops
__device__ __forceinline__ float sigmoid_sfu(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}


// sigmoid of the edge kernels: ACCURATE (fp32-accurate mode) = the two-MUFU form above; otherwise 0.5 + 0.5 tanh(x / 2) with
// one MUFU (tanh.approx: 2^-11 relative), well inside what bf16 operands (2^-9) leave of the 1e-2 budget of bf16 mode.
// The forward kernel is bound by the SFU (three sigmoid passes per element), so this is a third of its time in that mode.
template <bool ACCURATE>
__device__ __forceinline__ float sigmoid_t(float x) {
    if (ACCURATE) return sigmoid_sfu(x);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}

// ---- per-thread helpers of the edge-kernel epilogues ---------------------------------------------------------
// sum over the 32 lanes of 16 per-lane values by recursive halving (16 shuffles): lane l receives column (l >> 1)
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) {          // lane bit `off` selects which half of the remaining values it keeps
        const bool up = lane & off;
        const int h = off >> 1;                        // values kept after this step
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float send = up ? v[i] : v[i + h];
            const float keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Row sums of one thread's 16 consecutive edges as run partials (segment.cu): `run0` = run id of the first edge (< 0: the
// group holds no valid edge), `bits` = row-start flags, `nvalid` = how many of the 16 edges exist.  The header is
// warp-uniform, so the three cases are real branches: no row start inside the group (one sum), exactly one (two
// predicated sums), anything else or a ragged group (running sum flushed at every start).  The per-edge predicated
// flush alone cost 12 instructions per element.
__device__ __forceinline__ void run_sums16(const float (&x)[16], int run0, unsigned bits, int nvalid, float* __restrict__ runs,
                                           int n) {
    const unsigned inner = bits & 0xfffeu;
    float* rp = runs + (int64_t)run0 * 128 + n;
    if (nvalid >= 16 && inner == 0) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = x[2 * j] + x[2 * j + 1];
        *rp = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    } else if (nvalid >= 16 && (inner & (inner - 1)) == 0) {
        const int js = __ffs(inner) - 1;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j < js) a0 += x[j];
            else a1 += x[j];
        }
        rp[0] = a0;
        rp[128] = a1;
    } else if (run0 >= 0) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (j > 0 && ((inner >> j) & 1)) {
                *rp = acc;
                rp += 128;
                acc = 0.f;
            }
            acc += j < nvalid ? x[j] : 0.f;
        }
        *rp = acc;
    }
}

template <bool SPLIT, int KS, typename OA, typename OB>
__device__ __forceinline__ void issue_gemm_t(uint32_t tmem_d, uint64_t a, uint32_t a_lo_bytes, uint64_t b,
                                             uint32_t b_lo_bytes, uint32_t idesc, bool acc_first) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
        mma_f16(tmem_d, desc_at(a, OA::off(ks)), desc_at(b, OB::off(ks)), idesc, acc_first || ks > 0);
    if (SPLIT) {
        const uint64_t al = desc_at(a, a_lo_bytes), bl = desc_at(b, b_lo_bytes);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_f16(tmem_d, desc_at(a, OA::off(ks)), desc_at(bl, OB::off(ks)), idesc, true);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) mma_f16(tmem_d, desc_at(al, OA::off(ks)), desc_at(b, OB::off(ks)), idesc, true);
    }
}

}  // namespace tc
