// render_bwd_tma.cu — backward blend with a TMA (cp.async.bulk) + mbarrier multi-stage pipeline.
//
// Same per-pair mathematics as render_bwd.cu (shared text below), different data movement: instead of
// "all 256 threads gather a batch with LDG->STS, __syncthreads, compute, __syncthreads" the CTA has a
// ninth PRODUCER warp that, for every stage of 64 tile-list entries, reads the splat ids and issues
// one 96-byte bulk copy (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes — TMA, no
// register staging) per record into a ring of kStages shared-memory stages; completion is tracked by
// an mbarrier transaction count.  The eight consumer warps wait on the stage's "full" barrier,
// process it at their own pace and release it through an "empty" barrier.  There is no CTA-wide
// barrier in the loop, so a warp whose 8x4 footprint has little work runs ahead instead of idling at
// __syncthreads (ncu: 15 % of warp samples sat at the batch barriers in the classic kernel).
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

#ifndef SURFEL_BWD_BLOCKS
#define SURFEL_BWD_BLOCKS 4
#endif
constexpr int kStageSlots = 64;
constexpr int kStages = 3;
constexpr int kPanelRowT = 28;

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(288, SURFEL_BWD_BLOCKS) render_bwd_tma_kernel(RenderParams p) {
    __shared__ __align__(128) unsigned char s_stage[kStages * kStageSlots * kRecBytes];   // AoS records
    __shared__ __align__(16) float s_panel[8 * 32 * kPanelRowT];
    __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
    __shared__ uint32_t s_max[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;       // warp 8 = producer
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    const uint2 range = p.ranges[ty * p.gx + tx];
    const uint32_t stage0 = smem_u32(s_stage), full0 = smem_u32(s_full), empty0 = smem_u32(s_empty);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    int lx = 0, ly = 0;
    if (warp < 8) warp_pixel(warp, lane, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    const bool inside = warp < 8 && px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)(tx * kBlockX + ((warp & 1) << 3)), fx1 = fx0 + 7.0f;
    const float fy0 = (float)(ty * kBlockY + ((warp >> 1) << 2)), fy1 = fy0 + 3.0f;
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;
    const uint32_t panel_base = smem_u32(s_panel) + (warp & 7) * (32 * kPanelRowT * 4);
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr int kPanelRow = kPanelRowT;

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
    const uint32_t median_index = median_contributor - 1u;

    uint32_t warp_max = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));
    if (warp < 8 && lane == 0) s_max[warp] = warp_max;
    __syncthreads();                                   // also publishes the mbarrier inits
    uint32_t cta_max = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) cta_max = max(cta_max, s_max[w]);
    const int n_iter = ((int)cta_max + kStageSlots - 1) / kStageSlots;

    if (warp == 8) {
        // ===== producer: ids -> one TMA bulk copy per record, back to front =====
        for (int it = 0; it < n_iter; it++) {
            const int s = it % kStages;
            const int end = (int)cta_max - it * kStageSlots;
            const int n = min(kStageSlots, end), start = end - n;
            mbar_wait(empty0 + 8 * s, ((it / kStages) & 1) ^ 1);       // stage released by all 8 warps
            if (p.slab) {
                // sorted slab written by the forward: ONE bulk copy per stage
                if (lane == 0) {
                    mbar_expect_tx(full0 + 8 * s, (uint32_t)n * kRecBytes);
                    tma_bulk_g2s(stage0 + s * kStageSlots * kRecBytes, p.slab + (size_t)(range.x + start) * kRecQuads,
                                 (uint32_t)n * kRecBytes, full0 + 8 * s);
                }
            } else {
                uint32_t id0 = 0, id1 = 0;
                if (lane < n) id0 = p.point_list[range.x + start + lane];
                if (lane + 32 < n) id1 = p.point_list[range.x + start + lane + 32];
                if (lane == 0) mbar_expect_tx(full0 + 8 * s, (uint32_t)n * kRecBytes);
                __syncwarp();
                const uint32_t dst = stage0 + (s * kStageSlots + lane) * kRecBytes;
                if (lane < n) tma_bulk_g2s(dst, p.rec + (size_t)id0 * kRecQuads, kRecBytes, full0 + 8 * s);
                if (lane + 32 < n) tma_bulk_g2s(dst + 32 * kRecBytes, p.rec + (size_t)id1 * kRecQuads, kRecBytes, full0 + 8 * s);
            }
        }
        return;
    }

    // ===== consumers =====
    float T = T_final, S = 0.0f;
    constexpr float kMScale = kFar / (kFar - kNear);
    constexpr float kDmScale = (kFar * kNear) / (kFar - kNear);
    for (int it = 0; it < n_iter; it++) {
        const int s = it % kStages;
        const int end = (int)cta_max - it * kStageSlots;
        const int n = min(kStageSlots, end), start = end - n;
        const uint32_t stage_base = stage0 + s * kStageSlots * kRecBytes;
        mbar_wait(full0 + 8 * s, (it / kStages) & 1);
        if ((int)warp_max > start) {
            for (int c = ((n - 1) >> 5) << 5; c >= 0; c -= 32) {
                const int slot = c + lane;
                bool hit = false;
                if (slot < n && (uint32_t)(start + slot) < warp_max) {
                    const float4 bb = lds128(stage_base + slot * kRecBytes + 80);
                    hit = bb.x <= fx1 && bb.z >= fx0 && bb.y <= fy1 && bb.w >= fy0;
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int j = 31 - __clz(m);
                    m &= ~(1u << j);
                    const int k = c + j;
                    const uint32_t index = (uint32_t)(start + k);
                    const uint32_t ra = stage_base + k * kRecBytes;
                    const float4 q0 = lds128(ra), q1 = lds128(ra + 16), q2 = lds128(ra + 32);
                    PairEval e;
                    const bool active = index < last_contributor && eval_pair(pxf, pyf, q0, q1, q2, e);
                    const unsigned am = __ballot_sync(0xffffffffu, active);
                    if (am == 0u) continue;

                    if (active) {
                        const float4 q3 = lds128(ra + 48), q4 = lds128(ra + 64);
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
                        float dL_dz = (index == median_index) ? dL_dmedian : 0.0f;
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
                            const uint32_t id = __float_as_uint(lds32(ra + 64 + 12));
                            atomicAdd(p.grad_rec + (size_t)id * kGradFloats + lane, acc);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
}

int launch_render_bwd_tma(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderBwd, stream);
    render_bwd_tma_kernel<<<grid, 288, 0, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
