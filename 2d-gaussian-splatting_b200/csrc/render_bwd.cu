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
//  * per (warp, splat) the lanes that really contribute (on average ~9 of 32 at 1 M splats / 1080p,
//    ncu profiles/r1) are compacted with a ballot: each writes its 21 partials as one row
//    into a per-warp shared-memory panel, then 21 lanes each add one COLUMN of the panel and issue
//    ONE red.global.add.f32 into the splat's 96-byte gradient record (layout: common.cuh; the
//    homography gradient is carried as the sums A, Bx, By, Z and finished in preprocess backward).  Work scales with the number
//    of contributing lanes (a 32-lane shuffle butterfly cost 108 instructions per splat regardless),
//    and global atomics drop from 18 per (pixel,splat) to 18 per (warp,splat), contiguous.
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

#ifndef SURFEL_BWD_BLOCKS
#define SURFEL_BWD_BLOCKS 4
#endif
#ifndef SURFEL_BWD_BATCH
#define SURFEL_BWD_BATCH 192
#endif
constexpr int kBatchB = SURFEL_BWD_BATCH;   // 18 KB of records + 28 KB panel fit the 48 KB static limit
constexpr int kPanelRow = 28;                     // 21 used floats, 7-quad stride (odd: conflict-free STS.128)

__global__ void __launch_bounds__(256, SURFEL_BWD_BLOCKS) render_bwd_kernel(RenderParams p) {
    __shared__ float4 s_rec[kRecQuads * kBatchB];            // [quad][slot]; quad 4 .w carries the splat id
    __shared__ __align__(16) float s_panel[8 * 32 * kPanelRow];
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
    const uint32_t rec_base = smem_u32(s_rec);
    const uint32_t panel_base = smem_u32(s_panel) + warp * (32 * kPanelRow * 4);
    const unsigned lt_mask = (1u << lane) - 1u;

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
    const float bgT = -T_final * ((__ldg(p.bg + 0) * dpix0 + __ldg(p.bg + 1) * dpix1) + __ldg(p.bg + 2) * dpix2);
    const uint32_t median_index = median_contributor - 1u;   // 0xFFFFFFFE when there is none

    // warp / CTA extent of the replay
    uint32_t warp_max = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));
    if (lane == 0) s_max[warp] = warp_max;
    __syncthreads();
    uint32_t cta_max = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) cta_max = max(cta_max, s_max[w]);

    // Per-pixel replay state.  All blended outputs are linear in the weights w_j = alpha_j*T_j, so
    // with the per-pair scalar "value"
    //     v_j = rgb_j.dL_dC + depth_j*dL_dD + dL_dA + n_j.dL_dN + dL_dweight_j      (dL/dw_j)
    // the A.4 suffix recurrences (accum_rec for colour, depth, alpha, normal and last_dL_dT) collapse
    // into ONE accumulator S = sum_{j>i} w_j v_j:   dL/dalpha_i = T_i v_i - S/(1-alpha_i) + bg term.
    float T = T_final, S = 0.0f;
    constexpr float kMScale = kFar / (kFar - kNear);
    constexpr float kDmScale = (kFar * kNear) / (kFar - kNear);

    for (int end = (int)cta_max; end > 0; end -= kBatchB) {
        const int n = min(kBatchB, end);
        const int start = end - n;
        __syncthreads();
        if (tid < n) {
            const uint32_t id = p.point_list[range.x + start + tid];
            const float4* r = p.rec + (size_t)id * kRecQuads;
#pragma unroll
            for (int q = 0; q < kRecQuads; q++) {
                float4 v = __ldg(r + q);
                if (q == 4) v.w = __uint_as_float(id);
                s_rec[q * kBatchB + tid] = v;
            }
        }
        __syncthreads();
        if ((int)warp_max <= start) continue;   // nothing of this batch reaches this warp's pixels

        for (int c = ((n - 1) >> 5) << 5; c >= 0; c -= 32) {
            const int slot = c + lane;
            bool hit = false;
            if (slot < n && (uint32_t)(start + slot) < warp_max) {
                const float4 bb = lds128(rec_base + (5 * kBatchB + slot) * 16);
                hit = bb.x <= fx1 && bb.z >= fx0 && bb.y <= fy1 && bb.w >= fy0;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            uint32_t gb = rec_base + (uint32_t)c * 16u;
            asm volatile("" : "+r"(gb));      // keep it in a register (else re-derived from SR_CgaCtaId per hit)
            const int own = (int)last_contributor - (start + c);      // bits below `own` are this pixel's
            const int med = (int)median_index - (start + c);
            while (m) {
                const uint32_t j = high_bit(m);                       // back to front
                m &= low_mask(j);
                const uint32_t ra = gb + j * 16u;
                const float4 q0 = lds128(ra), q1 = lds128(ra + kBatchB * 16), q2 = lds128(ra + 2 * kBatchB * 16);
                PairEval e;
                const bool active = (int)j < own && eval_pair(pxf, pyf, q0, q1, q2, e);
                const unsigned am = __ballot_sync(0xffffffffu, active);
                if (am == 0u) continue;

                if (active) {
                    const float4 q3 = lds128(ra + 3 * kBatchB * 16), q4 = lds128(ra + 4 * kBatchB * 16);
                    const float G = e.G, alpha = e.alpha;
                    const float one_m = 1.0f - alpha;
                    const float inv1ma = fast_rcp(one_m);
                    T = T * inv1ma;
                    const float w = alpha * T;
                    const float inv_d = fast_rcp(e.depth);
                    const float m_d = kMScale * (1.0f - kNear * inv_d);
                    const float dmd_dd = kDmScale * inv_d * inv_d;
                    const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                    float v = dL_dweight + dL_daccum;
                    v = fmaf(q4.x, dpix0, v); v = fmaf(q4.y, dpix1, v); v = fmaf(q4.z, dpix2, v);
                    v = fmaf(e.depth, dL_ddepth, v);
                    v = fmaf(q3.x, dN0, v); v = fmaf(q3.y, dN1, v); v = fmaf(q3.z, dN2, v);
                    const float dL_dalpha = T * v - (S - bgT) * inv1ma;
                    S = fmaf(w, v, S);
                    float dL_dz = ((int)j == med) ? dL_dmedian : 0.0f;
                    dL_dz += 2.0f * w * (m_d * final_A - final_D) * dL_dreg * dmd_dd;
                    const float dL_dG = q2.w * dL_dalpha;
                    dL_dz += w * dL_ddepth;
                    float ax = 0, ay = 0, az = 0, zx = 0, zy = 0, m2x = 0, m2y = 0;
                    if (e.use3d) {
                        const float Twx = q1.z, Twy = q1.w;
                        const float nG = -G * dL_dG;
                        const float dsx = nG * e.sx + dL_dz * Twx;
                        const float dsy = nG * e.sy + dL_dz * Twy;
                        ax = dsx * e.inv_pz; ay = dsy * e.inv_pz;
                        az = -(ax * e.sx + ay * e.sy);
                        zx = dL_dz * e.sx; zy = dL_dz * e.sy;
                    } else {
                        const float gg = -G * kFilterInvSquare * dL_dG;
                        m2x = gg * e.dx; m2y = gg * e.dy;
                        if (p.lowpass_quirk) { zx = e.sx * dL_dz; zy = e.sy * dL_dz; }
                    }
                    const float ndx = -e.dx, ndy = -e.dy;      // pixel - AABB centre
                    // one row per contributing lane (rows are compacted: ballot prefix)
                    const uint32_t row = panel_base + __popc(am & lt_mask) * (kPanelRow * 4);
                    sts128(row, make_float4(ax, ay, az, ndx * ax));
                    sts128(row + 16, make_float4(ndx * ay, ndx * az, ndy * ax, ndy * ay));
                    sts128(row + 32, make_float4(ndy * az, zx, zy, dL_dz));
                    sts128(row + 48, make_float4(m2x, m2y, G * dL_dalpha, w * dN0));
                    sts128(row + 64, make_float4(w * dN1, w * dN2, w * dpix0, w * dpix1));
                    sts32(row + 80, w * dpix2);
                }
                __syncwarp();
                if (lane < kGradUsed) {
                    const int nact = __popc(am);
                    uint32_t a = panel_base + lane * 4;
                    float acc = 0.0f;
                    for (int r = 0; r < nact; r++, a += kPanelRow * 4) acc += lds32(a);
                    if (acc != 0.0f) {
                        const uint32_t id = __float_as_uint(lds32(ra + 4 * kBatchB * 16 + 12));
                        atomicAdd(p.grad_rec + (size_t)id * kGradFloats + lane, acc);
                    }
                }
                __syncwarp();
            }
        }
    }
}

int launch_render_bwd(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderBwd, stream);
    render_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
