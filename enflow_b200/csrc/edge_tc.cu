// Weight images of the tensor-core kernels: fp32 parameters -> swizzled bf16 operand images (hi and lo of the bf16x3
// split), all layers in two launches; and the entry point of the forward edge kernel (edge_tc_fwd.cu).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

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

// mode 1 = split (fp32-accurate), mode 2 = bf16; the kernel is edge_tc_fwd.cu
int enf_edge_fwd_tc(int mode, const int* row, const int* col, const int* E_dev, int E_cap, const float* pos,
                    const float* box, const float* P, const float* S, const float* lp, const unsigned char* wimg,
                    int nf, const int* rowptr, const int* mis, float* runs, float* s_out, float* trans,
                    cudaStream_t st) {
    return enf_edge_fwd_tc_v2(mode, row, col, E_dev, E_cap, pos, box, P, S, lp, wimg, nf, rowptr, mis, runs, s_out, trans, st);
}
