// Fused Adam on the flat parameter buffer (the step right after the hot path,
// enflow/main.py:177,222).  One kernel updates all 269k parameters in place from the flat gradient buffer the
// backward pass (and the data-parallel all-reduce) just produced; torch.optim.Adam semantics (no amsgrad, no
// weight decay).  The step counter lives on the device so the launch can be captured in a CUDA graph.
#include "common.cuh"

namespace {

__global__ void k_adam_tick(int* __restrict__ step) { step[0] += 1; }

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, int64_t n, const int* __restrict__ step, float lr,
                                               const float* __restrict__ lr_dev, float beta1, float beta2, float eps) {
    if (lr_dev) lr = lr_dev[0];          // device-resident learning rate: a captured graph follows a scheduler
    const float t = (float)step[0];
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = fmaf(beta1, m[i], (1.0f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

}  // namespace

int enf_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int* step, float lr, const float* lr_dev,
                  float beta1, float beta2, float eps, cudaStream_t st) {
    if (n == 0) return ENF_OK;
    enf_count_launch(), k_adam_tick<<<1, 1, 0, st>>>(step);
    int blocks = (int)((n + 255) / 256);
    if (blocks > enf_num_sms() * 8) blocks = enf_num_sms() * 8;
    enf_count_launch(), k_adam<<<blocks, 256, 0, st>>>(p, g, m, v, n, step, lr, lr_dev, beta1, beta2, eps);
    ENF_CHECK_LAUNCH();
    return ENF_OK;
}
