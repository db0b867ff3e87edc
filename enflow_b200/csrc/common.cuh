// Shared device helpers and the flat parameter layout for the enflow B200 path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ENF_H 128          // hidden width the kernels are specialised for (example/train.yaml:19)
#define ENF_MAX_NF 8       // node features: 1 (LJ), 4 (train.yaml shape), 5 (atom_types, constants.py:7)
#define ENF_TILE_E 128     // edges per tile (one MMA M-block)
#define ENF_ALIGN 32       // every parameter tensor starts on a 32-float (128 B) boundary

#define ENF_OK 0
#define ENF_ERR_ARG 1
#define ENF_ERR_CUDA 2

void enf_set_error(const char* fmt, ...);
void enf_count_launch();                                  // every kernel launch of this library is counted
// optional per-kernel-family timing with CUDA events on the launch stream (see enflow_timing_*)
// one family per kernel (a family that mixes kernels cannot be held against a roofline)
enum { TK_EDGES = 0, TK_NODE_PRE, TK_EDGE_FWD, TK_RUN_SUM, TK_SEG_COLS, TK_SEG_ROWS, TK_SEG3, TK_NODE_POST, TK_COUPLING_FWD,
       TK_COUPLING_BWD, TK_COUPLING_INV, TK_EDGE_GEOM, TK_EDGE_BWD, TK_EDGE_REDUCE, TK_NODE_POST_BWD, TK_NODE_PRE_BWD,
       TK_COL_PERM, TK_ARGMAX, TK_NLL, TK_COUNT };
void enf_time_begin(int kind, cudaStream_t st);
bool enf_timing_on();
// `to` waits for everything enqueued on `from` so far (event record + stream wait; capturable); no-op when from == to
void enf_chain(cudaStream_t from, cudaStream_t to);
// library-owned side streams (created once per process = once per device); [0] carries the weight-gradient reductions,
// kernels that nothing on the main stream waits for until the end of the backward pass
cudaStream_t enf_side_stream(int which);
// named events (id 0..7) for dependencies that are not "everything so far": record on one stream, wait on another
void enf_mark(int id, cudaStream_t st);
void enf_wait_mark(int id, cudaStream_t st);
void enf_time_end(cudaStream_t st);

#define ENF_CHECK_ARG(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            enf_set_error(__VA_ARGS__);          \
            return ENF_ERR_ARG;                  \
        }                                        \
    } while (0)

#define ENF_CHECK_LAUNCH()                                                        \
    do {                                                                          \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            enf_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return ENF_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

#define ENF_TRY(call)               \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != ENF_OK) return rc_; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Flat parameter buffer. One EGCL layer (reference state_dict names, enflow/nn/egcl.py:21-55):
//   W1 edge_nn.0.weight [H,2nf+1]  b1 edge_nn.0.bias [H]   W2 edge_nn.2.weight [H,H]  b2 [H]
//   W4 node_nn.0.weight [H,H+nf]   b4 [H]                   W5 node_nn.2.weight [nf,H] b5 [nf]
//   W3 coord_nn.0.weight [H,H]     b3 [H]                   wc coord_nn.2.weight [1,H]
//   W6 vel_scaling_nn.0.weight [H,nf] b6 [H]                W7 vel_scaling_nn.2.weight [1,H] b7 [1]
// ArgMax (enflow/nn/argmax.py:9-12): Wa0 [H,nf] ba0 [H] Wa2 [2nf,H] ba2 [2nf].
// Gradients use the same layout in a second flat buffer (one NCCL all-reduce covers everything).
enum { P_W1 = 0, P_B1, P_W2, P_B2, P_W4, P_B4, P_W5, P_B5, P_W3, P_B3, P_WC, P_W6, P_B6, P_W7, P_B7, P_EGCL_COUNT };
enum { PA_W0 = 0, PA_B0, PA_W2, PA_B2, PA_COUNT };

struct EgclOffsets {
    int64_t off[P_EGCL_COUNT];
    int64_t size;   // padded floats per layer
};

__host__ __device__ inline int64_t enf_pad(int64_t n) { return (n + ENF_ALIGN - 1) / ENF_ALIGN * ENF_ALIGN; }

inline void enf_egcl_sizes(int nf, int64_t* sz) {
    const int64_t H = ENF_H;
    sz[P_W1] = H * (2 * nf + 1); sz[P_B1] = H; sz[P_W2] = H * H; sz[P_B2] = H;
    sz[P_W4] = H * (H + nf); sz[P_B4] = H; sz[P_W5] = (int64_t)nf * H; sz[P_B5] = nf;
    sz[P_W3] = H * H; sz[P_B3] = H; sz[P_WC] = H;
    sz[P_W6] = H * nf; sz[P_B6] = H; sz[P_W7] = H; sz[P_B7] = 1;
}

inline EgclOffsets enf_egcl_offsets(int nf) {
    EgclOffsets o;
    int64_t sz[P_EGCL_COUNT];
    enf_egcl_sizes(nf, sz);
    int64_t cur = 0;
    for (int i = 0; i < P_EGCL_COUNT; ++i) { o.off[i] = cur; cur += enf_pad(sz[i]); }
    o.size = cur;
    return o;
}

inline void enf_argmax_sizes(int nf, int64_t* sz) {
    const int64_t H = ENF_H;
    sz[PA_W0] = H * nf; sz[PA_B0] = H; sz[PA_W2] = 2 * (int64_t)nf * H; sz[PA_B2] = 2 * nf;
}

struct ArgmaxOffsets {
    int64_t off[PA_COUNT];
    int64_t size;
};

inline ArgmaxOffsets enf_argmax_offsets(int nf) {
    ArgmaxOffsets o;
    int64_t sz[PA_COUNT];
    enf_argmax_sizes(nf, sz);
    int64_t cur = 0;
    for (int i = 0; i < PA_COUNT; ++i) { o.off[i] = cur; cur += enf_pad(sz[i]); }
    o.size = cur;
    return o;
}

// Per-layer packed (derived) weights, rebuilt once per forward by enflow_pack_weights:
//   W2T [H][H], W3T [H][H] (k-major rows for the forward GEMMs), W4T [H+nf][H], W4A [H][H] = W4[:, nf:].
struct PackOffsets {
    int64_t w2t, w3t, w4t, w4a, size;
};
inline PackOffsets enf_pack_offsets(int nf) {
    PackOffsets p;
    const int64_t H = ENF_H;
    p.w2t = 0; p.w3t = H * H; p.w4t = 2 * H * H; p.w4a = enf_pad(2 * H * H + (H + nf) * H);
    p.size = p.w4a + H * H;
    return p;
}

// ---------------------------------------------------------------------------------------------
// fp32 mode keeps the accurate expf/division (parity budget 1e-5 end to end); the bf16 tensor
// path uses the fast forms below.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float sigmoid_fast_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// d/dx silu(x) = s (1 + x (1 - s))
__device__ __forceinline__ float dsiluf_(float x) {
    float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x - rint(x/p)*p, torch.round is half-to-even (enflow/utils/helpers.py:7-8)
__device__ __forceinline__ float wrapf_(float x, float p) { return x - rintf(x / p) * p; }

// Sum of per-CTA partials for one output element, shared by the *_reduce kernels.  Launch with 256 threads and
// ceil(stride / 32) CTAs: a block handles 32 consecutive elements x 8 CTA groups; group y adds CTAs y, y+8, ... in
// order (coalesced 128-byte reads), the eight group sums are then added in order by the first warp.  Deterministic.
// Returns true (with idx, acc set) for the one thread per element that must store the result.
__device__ __forceinline__ bool enf_reduce_partials_32x8(const float* __restrict__ partial, int n_cta, int64_t stride,
                                                         int& idx, float& acc) {
    __shared__ float part_s[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    idx = blockIdx.x * 32 + tx;
    float a0 = 0.f, a1 = 0.f;
    if (idx < stride) {
        int c = ty;
        for (; c + 8 < n_cta; c += 16) {
            a0 += partial[(int64_t)c * stride + idx];
            a1 += partial[(int64_t)(c + 8) * stride + idx];
        }
        if (c < n_cta) a0 += partial[(int64_t)c * stride + idx];
    }
    part_s[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty != 0 || idx >= stride) return false;
    acc = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) acc += part_s[y][tx];
    return true;
}

static inline int enf_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
