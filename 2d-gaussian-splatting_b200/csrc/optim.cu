// optim.cu — SURVEY §8(f) row f3: the parameter update that follows the rasterizer backward.
//
//  * surfel_adam_step: ONE launch applies Adam to every parameter group of the Gaussian model
//    (reference: torch.optim.Adam(l, lr=0.0, eps=1e-15) built at /root/reference/scene/gaussian_model.py:148-166
//    and stepped at /root/reference/train.py:138-140).  The reference's default (foreach) Adam
//    makes ~12 elementwise passes per group over param / grad / exp_avg / exp_avg_sq; here each
//    element is read and written once: 7 floats of traffic per parameter (grad r, param rw, m rw, v rw),
//    i.e. 59 parameters x 28 B = 1652 B per splat — pure HBM streaming, float4-vectorised.
//    Arithmetic follows torch's single-tensor Adam op for op (lerp, addcmul, sqrt / bias2 + eps,
//    addcdiv) so that results agree to float rounding.
//  * surfel_densify_stats: the three in-place statistics of /root/reference/train.py:125-128 +
//    /root/reference/scene/gaussian_model.py:405-407 in one pass: max_radii2D = max(., radii),
//    xyz_gradient_accum += |means2D.grad|, denom += 1, all only where radii > 0.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/surfel_rasterizer.h"
#include "common.cuh"
#include "profile.h"

namespace surfel {

struct AdamTable {
    surfel_adam_group_t g[SURFEL_ADAM_MAX_GROUPS];
    long long first_block[SURFEL_ADAM_MAX_GROUPS + 1];   // block index where each group starts
    int n;
    float beta2, w1, w2, eps;   // w = 1 - beta, rounded from double like torch's scalar arguments
};

constexpr int kAdamThreads = 256;
constexpr int kAdamPerThread = 4;                          // one float4 per array per thread
constexpr int kAdamPerBlock = kAdamThreads * kAdamPerThread;

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float w1, float beta2, float w2,
                                         float step_size, float bc2_sqrt, float eps) {
    m = m + w1 * (g - m);                                  // exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + w2 * g * g;                            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + eps;       // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - step_size * (m / denom);                       // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const __grid_constant__ AdamTable t) {
    // which group does this block belong to (<= 8 groups: linear search on block-uniform data)
    int gi = 0;
    const long long b = blockIdx.x;
#pragma unroll
    for (int i = 1; i < SURFEL_ADAM_MAX_GROUPS; i++)
        if (i < t.n && b >= t.first_block[i]) gi = i;
    const surfel_adam_group_t& G = t.g[gi];
    const long long base = (b - t.first_block[gi]) * kAdamPerBlock + (long long)threadIdx.x * kAdamPerThread;
    if (base >= G.n) return;
    const float w1 = t.w1, w2 = t.w2;
    const bool vec = (G.n - base) >= kAdamPerThread && G.aligned16;
    if (vec) {
        float4 p = *(float4*)(G.param + base);
        const float4 g = __ldg((const float4*)(G.grad + base));
        float4 m = *(float4*)(G.exp_avg + base), v = *(float4*)(G.exp_avg_sq + base);
        adam_one(p.x, g.x, m.x, v.x, w1, t.beta2, w2, G.step_size, G.bias2_sqrt, t.eps);
        adam_one(p.y, g.y, m.y, v.y, w1, t.beta2, w2, G.step_size, G.bias2_sqrt, t.eps);
        adam_one(p.z, g.z, m.z, v.z, w1, t.beta2, w2, G.step_size, G.bias2_sqrt, t.eps);
        adam_one(p.w, g.w, m.w, v.w, w1, t.beta2, w2, G.step_size, G.bias2_sqrt, t.eps);
        *(float4*)(G.param + base) = p;
        *(float4*)(G.exp_avg + base) = m;
        *(float4*)(G.exp_avg_sq + base) = v;
    } else {
        for (long long i = base; i < G.n && i < base + kAdamPerThread; i++) {
            float p = G.param[i], m = G.exp_avg[i], v = G.exp_avg_sq[i];
            adam_one(p, G.grad[i], m, v, w1, t.beta2, w2, G.step_size, G.bias2_sqrt, t.eps);
            G.param[i] = p; G.exp_avg[i] = m; G.exp_avg_sq[i] = v;
        }
    }
}

__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const int32_t* __restrict__ radii,
                                                            const float* __restrict__ grad2d, float* accum,
                                                            float* denom, float* max_radii) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = grad2d[3 * (size_t)i + 0], gy = grad2d[3 * (size_t)i + 1], gz = grad2d[3 * (size_t)i + 2];
    accum[i] += sqrtf(gx * gx + gy * gy + gz * gz);        // torch.norm(grad[filter], dim=-1)
    denom[i] += 1.0f;
    if (max_radii) max_radii[i] = fmaxf(max_radii[i], (float)r);
}

}  // namespace surfel

using namespace surfel;

extern "C" {

int surfel_adam_step(int n_groups, const surfel_adam_group_t* groups, double beta1, double beta2, double eps,
                     void* stream) {
    if (n_groups < 0 || n_groups > SURFEL_ADAM_MAX_GROUPS) {
        surfel_set_error("surfel_adam_step: n_groups %d outside [0, %d]", n_groups, SURFEL_ADAM_MAX_GROUPS);
        return 1;
    }
    if (n_groups && !groups) { surfel_set_error("surfel_adam_step: NULL groups"); return 1; }
    AdamTable t;
    t.n = 0; t.beta2 = (float)beta2; t.w1 = (float)(1.0 - beta1); t.w2 = (float)(1.0 - beta2); t.eps = (float)eps;
    long long blocks = 0;
    for (int i = 0; i < n_groups; i++) {
        const surfel_adam_group_t& g = groups[i];
        if (g.n < 0) { surfel_set_error("surfel_adam_step: group %d has n < 0", i); return 1; }
        if (g.n == 0) continue;
        if (!g.param || !g.grad || !g.exp_avg || !g.exp_avg_sq) {
            surfel_set_error("surfel_adam_step: group %d has a NULL pointer", i);
            return 1;
        }
        t.g[t.n] = g;
        t.g[t.n].aligned16 = ((((uintptr_t)g.param | (uintptr_t)g.grad | (uintptr_t)g.exp_avg | (uintptr_t)g.exp_avg_sq) & 15) == 0);
        t.first_block[t.n] = blocks;
        blocks += (g.n + kAdamPerBlock - 1) / kAdamPerBlock;
        t.n++;
    }
    t.first_block[t.n] = blocks;
    if (blocks == 0) return 0;
    if (blocks > 0x7fffffffLL) { surfel_set_error("surfel_adam_step: too many elements"); return 1; }
    LaunchScope scope(kStAdam, (cudaStream_t)stream);
    adam_kernel<<<(unsigned)blocks, kAdamThreads, 0, (cudaStream_t)stream>>>(t);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

int surfel_densify_stats(int P, const int32_t* radii, const float* means2D_grad, float* xyz_gradient_accum,
                         float* denom, float* max_radii2D, void* stream) {
    if (P < 0) { surfel_set_error("surfel_densify_stats: P < 0"); return 1; }
    if (P == 0) return 0;
    if (!radii || !means2D_grad || !xyz_gradient_accum || !denom) {
        surfel_set_error("surfel_densify_stats: NULL required pointer");
        return 1;
    }
    LaunchScope scope(kStDensifyStats, (cudaStream_t)stream);
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, radii, means2D_grad, xyz_gradient_accum,
                                                                           denom, max_radii2D);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
