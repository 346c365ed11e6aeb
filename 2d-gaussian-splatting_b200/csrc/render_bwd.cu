// render_bwd.cu — per-tile back-to-front replay, backward of the fused blend.
//
// Replaces upstream renderCUDA backward (algorithm SURVEY Appendix A.4): for each
// pixel walk the tile list from the pixel's last contributor to the front, rebuild T by division,
// and accumulate the gradient of all 10 output channels (RGB, expected depth, alpha, normal,
// median depth, distortion) into per-splat dL_dtransMat[9], dL_dmean2D[2], dL_dopacity,
// dL_dnormal[3], dL_dcolor[3].
//
// B200 design (not upstream's, which issues up to 16 global float atomics per (pixel,splat)):
//  * same 96-byte records / 8x4 warp footprints / bbox ballot culling as the forward (exact);
//  * the CTA starts at the largest last_contributor of its pixels, not at the end of the list;
//  * the 18 per-pair partials are reduced ACROSS THE WARP first: a 16-value halving butterfly
//    (16 shuffles) + 2 plain xor-reductions, after which 18 lanes issue one RED each into the
//    splat's 80-byte gradient record — 18 contiguous atomics per (warp,splat) instead of 18 per
//    (pixel,splat), and none at all when no lane of the warp got a contribution.
#include "render_common.cuh"
#include "kernels.h"

namespace surfel {

constexpr int kBatchB = 256;

// Sum v[0..15] over the 32 lanes; on return lane L holds the total of value index (L >> 1) in v[0].
__device__ __forceinline__ void warp_reduce16(float (&v)[16], int lane) {
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float send = up ? v[j] : v[j + 8];
            const float keep = up ? v[j + 8] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float send = up ? v[j] : v[j + 4];
            const float keep = up ? v[j + 4] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float send = up ? v[j] : v[j + 2];
            const float keep = up ? v[j + 2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = lane & 2;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__global__ void __launch_bounds__(256, 2) render_bwd_kernel(RenderParams p) {
    __shared__ float4 s_rec[kRecQuads * kBatchB];   // [quad][slot]
    __shared__ uint32_t s_id[kBatchB];
    __shared__ uint32_t s_max[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    int lx, ly;
    warp_pixel(warp, lane, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)(tx * kBlockX + ((warp & 1) << 3)), fx1 = fx0 + 7.0f;
    const float fy0 = (float)(ty * kBlockY + ((warp >> 1) << 2)), fy1 = fy0 + 3.0f;
    const uint2 range = p.ranges[ty * p.gx + tx];
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;

    float T_final = 0, final_D = 0, final_D2 = 0;
    uint32_t last_contributor = 0, median_contributor = 0;
    float dpix0 = 0, dpix1 = 0, dpix2 = 0, dN0 = 0, dN1 = 0, dN2 = 0;
    float dL_ddepth = 0, dL_daccum = 0, dL_dreg = 0, dL_dmedian = 0;
    if (inside) {
        T_final = p.accum[pix]; final_D = p.accum[HW + pix]; final_D2 = p.accum[2 * HW + pix];
        last_contributor = p.n_contrib[pix]; median_contributor = p.n_contrib[HW + pix];
        dpix0 = p.dL_dpix[pix]; dpix1 = p.dL_dpix[HW + pix]; dpix2 = p.dL_dpix[2 * HW + pix];
        dL_ddepth = p.dL_dothers[kChDepth * HW + pix];
        dL_daccum = p.dL_dothers[kChAlpha * HW + pix];
        dN0 = p.dL_dothers[(kChNormal + 0) * HW + pix];
        dN1 = p.dL_dothers[(kChNormal + 1) * HW + pix];
        dN2 = p.dL_dothers[(kChNormal + 2) * HW + pix];
        dL_dmedian = p.dL_dothers[kChMidDepth * HW + pix];
        dL_dreg = p.dL_dothers[kChDistortion * HW + pix];
    }
    const float final_A = 1.0f - T_final;
    const float bg_dot = (__ldg(p.bg + 0) * dpix0 + __ldg(p.bg + 1) * dpix1) + __ldg(p.bg + 2) * dpix2;

    // warp / CTA extent of the replay
    uint32_t warp_max = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));
    if (lane == 0) s_max[warp] = warp_max;
    __syncthreads();
    uint32_t cta_max = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) cta_max = max(cta_max, s_max[w]);

    float T = T_final, last_alpha = 0, last_dL_dT = 0;
    float last_c0 = 0, last_c1 = 0, last_c2 = 0, acc_c0 = 0, acc_c1 = 0, acc_c2 = 0;
    float last_depth = 0, acc_depth = 0, acc_alpha = 0;
    float last_n0 = 0, last_n1 = 0, last_n2 = 0, acc_n0 = 0, acc_n1 = 0, acc_n2 = 0;

    for (int end = (int)cta_max; end > 0; end -= kBatchB) {
        const int n = min(kBatchB, end);
        const int start = end - n;
        __syncthreads();
        if (tid < n) {
            const uint32_t id = p.point_list[range.x + start + tid];
            s_id[tid] = id;
            const float4* r = p.rec + (size_t)id * kRecQuads;
#pragma unroll
            for (int q = 0; q < kRecQuads; q++) s_rec[q * kBatchB + tid] = __ldg(r + q);
        }
        __syncthreads();
        if ((int)warp_max <= start) continue;   // nothing of this batch reaches this warp's pixels

        for (int c = ((n - 1) >> 5) << 5; c >= 0; c -= 32) {
            const int slot = c + lane;
            bool hit = false;
            if (slot < n && (uint32_t)(start + slot) < warp_max) {
                const float4 bb = s_rec[5 * kBatchB + slot];
                hit = bb.x <= fx1 && bb.z >= fx0 && bb.y <= fy1 && bb.w >= fy0;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int j = 31 - __clz(m);
                m &= ~(1u << j);
                const int k = c + j;
                const uint32_t index = (uint32_t)(start + k);   // 0-based contributor
                const float4 q0 = s_rec[0 * kBatchB + k], q1 = s_rec[1 * kBatchB + k], q2 = s_rec[2 * kBatchB + k];
                PairEval e;
                const bool active = index < last_contributor && eval_pair(pxf, pyf, q0, q1, q2, e);
                if (!__any_sync(0xffffffffu, active)) continue;

                float g[16];
#pragma unroll
                for (int i = 0; i < 16; i++) g[i] = 0.0f;
                float gc1 = 0.0f, gc2 = 0.0f;
                if (active) {
                    const float4 q3 = s_rec[3 * kBatchB + k], q4 = s_rec[4 * kBatchB + k];
                    const float G = e.G, alpha = e.alpha;
                    T = T / (1.0f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.0f;
                    // colour
                    acc_c0 = last_alpha * last_c0 + (1.0f - last_alpha) * acc_c0; last_c0 = q4.x;
                    acc_c1 = last_alpha * last_c1 + (1.0f - last_alpha) * acc_c1; last_c1 = q4.y;
                    acc_c2 = last_alpha * last_c2 + (1.0f - last_alpha) * acc_c2; last_c2 = q4.z;
                    dL_dalpha += (q4.x - acc_c0) * dpix0 + (q4.y - acc_c1) * dpix1 + (q4.z - acc_c2) * dpix2;
                    g[15] = w * dpix0; gc1 = w * dpix1; gc2 = w * dpix2;
                    // distortion + median depth
                    float dL_dz = 0.0f;
                    const float m_d = kFar / (kFar - kNear) * (1.0f - kNear / e.depth);
                    const float dmd_dd = (kFar * kNear) / ((kFar - kNear) * e.depth * e.depth);
                    if (index == median_contributor - 1u) dL_dz += dL_dmedian;
                    const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1.0f - alpha) * last_dL_dT;
                    dL_dz += 2.0f * w * (m_d * final_A - final_D) * dL_dreg * dmd_dd;
                    // expected depth, alpha
                    acc_depth = last_alpha * last_depth + (1.0f - last_alpha) * acc_depth; last_depth = e.depth;
                    dL_dalpha += (e.depth - acc_depth) * dL_ddepth;
                    acc_alpha = last_alpha + (1.0f - last_alpha) * acc_alpha;
                    dL_dalpha += (1.0f - acc_alpha) * dL_daccum;
                    // normal
                    acc_n0 = last_alpha * last_n0 + (1.0f - last_alpha) * acc_n0; last_n0 = q3.x;
                    acc_n1 = last_alpha * last_n1 + (1.0f - last_alpha) * acc_n1; last_n1 = q3.y;
                    acc_n2 = last_alpha * last_n2 + (1.0f - last_alpha) * acc_n2; last_n2 = q3.z;
                    dL_dalpha += (q3.x - acc_n0) * dN0 + (q3.y - acc_n1) * dN1 + (q3.z - acc_n2) * dN2;
                    g[12] = w * dN0; g[13] = w * dN1; g[14] = w * dN2;

                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = q2.w * dL_dalpha;
                    dL_dz += w * dL_ddepth;
                    if (e.use3d) {
                        const float Twx = q1.z, Twy = q1.w;
                        const float dsx = dL_dG * -G * e.sx + dL_dz * Twx;
                        const float dsy = dL_dG * -G * e.sy + dL_dz * Twy;
                        const float inv = fast_rcp(e.pz);
                        const float dpx = dsx * inv, dpy = dsy * inv;
                        const float dpz = -(dpx * e.sx + dpy * e.sy);
                        const float dkx = e.ly * dpz - e.lz * dpy, dky = e.lz * dpx - e.lx * dpz, dkz = e.lx * dpy - e.ly * dpx;
                        const float dlx = dpy * e.kz - dpz * e.ky, dly = dpz * e.kx - dpx * e.kz, dlz = dpx * e.ky - dpy * e.kx;
                        g[0] = -dkx; g[1] = -dky; g[2] = -dkz;
                        g[3] = -dlx; g[4] = -dly; g[5] = -dlz;
                        g[6] = pxf * dkx + pyf * dlx + dL_dz * e.sx;
                        g[7] = pxf * dky + pyf * dly + dL_dz * e.sy;
                        g[8] = pxf * dkz + pyf * dlz + dL_dz;
                    } else {
                        const float gg = -G * kFilterInvSquare * dL_dG;
                        g[9] = gg * e.dx; g[10] = gg * e.dy;
                        if (p.lowpass_quirk) { g[6] = e.sx * dL_dz; g[7] = e.sy * dL_dz; }
                        g[8] = dL_dz;
                    }
                    g[11] = G * dL_dalpha;
                }
                warp_reduce16(g, lane);
                gc1 = warp_sum(gc1);
                gc2 = warp_sum(gc2);
                float* dst = p.grad_rec + (size_t)s_id[k] * kGradFloats;
                if ((lane & 1) == 0) {
                    if (g[0] != 0.0f) atomicAdd(dst + (lane >> 1), g[0]);
                } else if (lane == 1) {
                    if (gc1 != 0.0f) atomicAdd(dst + 16, gc1);
                } else if (lane == 3) {
                    if (gc2 != 0.0f) atomicAdd(dst + 17, gc2);
                }
            }
        }
    }
}

int launch_render_bwd(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    dim3 grid(p.gx, rows);
    render_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
