// render_bwd.cu — per-tile back-to-front replay, backward of the fused blend.
//
// Replaces upstream renderCUDA backward (algorithm SURVEY Appendix A.4): for each
// pixel walk the tile list from the pixel's last contributor to the front, rebuild T by division,
// and accumulate the gradient of all 10 output channels (RGB, expected depth, alpha, normal,
// median depth, distortion) into per-splat dL_dtransMat[9], dL_dmean2D[2], dL_dopacity,
// dL_dnormal[3], dL_dcolor[3].
//
// B200 design (not upstream's, which issues up to 16 global float atomics per (pixel,splat)):
//  * same 128-byte records / staging-time footprint classification / 8x4 warp footprints / affine
//    ray-splat intersection as the forward (render_fwd.cu); the classification is exact, and the
//    per-pair decisions are bit-identical to the forward's (same inline evaluation);
//  * the CTA starts at the largest last_contributor of its pixels, not at the end of the list, and for
//    all but crowded tiles the whole replay is staged in one round (no block barrier in the hit loop);
//  * A.4's eight suffix recurrences are collapsed into ONE accumulator S (see below);
//  * per (warp, splat) the lanes that really contribute (~9 of 32 at 1 M splats / 1080p) are compacted
//    with a ballot: each writes its 22 partials as one row of a per-warp shared-memory panel, then 22
//    lanes each add one COLUMN of the panel and issue ONE red.global.add.f32 into the splat's 96-byte gradient
//    record (layout: common.cuh; the homography gradient is carried as the sums A, Bx, By, Z and finished
//    in preprocess backward).  Global atomics drop from 18 per (pixel,splat) to 18 per (warp,splat),
//    contiguous.
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

#ifndef SURFEL_BWD_BLOCKS
#define SURFEL_BWD_BLOCKS 4
#endif
#ifndef SURFEL_BWD_BATCH
#define SURFEL_BWD_BATCH 256
#endif
constexpr int kBatchB = SURFEL_BWD_BATCH;         // multiple of 32
constexpr int kGroupsB = kBatchB / 32;
constexpr int kPanelRow = 24;                     // 22 used floats.  Rows 4 apart share banks (two-way conflicts on the row stores when a
                                                  // half warp holds >= 5 live lanes); the conflict-free 28-float stride was measured SLOWER
                                                  // (0.882 vs 0.865 ms): the 4 KB of shared memory it costs per CTA matter more
constexpr int kPanelBytes = 8 * 32 * kPanelRow * 4;
constexpr int kRecBytesB = kRecQuadsFwd * kBatchB * 16;        // the backward stages the same five quads as the forward
constexpr int kBwdSmemBytes = kRecBytesB + kBatchB * 4 + kPanelBytes + 8 * kGroupsB * 4 + 32;

// Sum of the first `cnt` (1..32) rows of this lane's panel column: blocks of four rows (independent loads in
// flight, two partial sums), then the 0-3 leftover rows.  The loops are kept rolled: unrolled by the
// compiler they turn into a tree of trip-count tests that costs more than the rows themselves at ~9 rows
// (measured at the headline workload: 0.878 ms this way, 0.893 ms compiler-unrolled, 0.963 ms one row per
// iteration, 0.958 ms through an indexed-branch ladder — BRX is slow).
__device__ __forceinline__ float column_sum(uint32_t a, int cnt) {
    float acc = 0.0f, acc2 = 0.0f;
    int r = cnt;
#pragma unroll 1
    for (; r >= 4; r -= 4, a += 4 * kPanelRow * 4) {
        const float x0 = lds32(a), x1 = lds32(a + kPanelRow * 4), x2 = lds32(a + 2 * kPanelRow * 4), x3 = lds32(a + 3 * kPanelRow * 4);
        acc += x0 + x1; acc2 += x2 + x3;
    }
#pragma unroll 1
    for (; r > 0; r--, a += kPanelRow * 4) acc += lds32(a);
    return acc + acc2;
}

__global__ void __launch_bounds__(256, SURFEL_BWD_BLOCKS) render_bwd_kernel(RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_rec = reinterpret_cast<float4*>(smem_raw);                                     // [quad][slot]
    uint32_t* s_id = reinterpret_cast<uint32_t*>(smem_raw + kRecBytesB);                     // [slot] splat index
    float* s_panel = reinterpret_cast<float*>(smem_raw + kRecBytesB + kBatchB * 4);
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw + kRecBytesB + kBatchB * 4 + kPanelBytes);  // [warp][group]
    uint32_t* s_max = s_mask + 8 * kGroupsB;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    int lx, ly;
    warp_pixel(warp, lane, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tx * kBlockX), oy = (float)(ty * kBlockY);
    const uint2 range = p.ranges[ty * p.gx + tx];
    const size_t HW = (size_t)p.H * p.W;
    const size_t pix = (size_t)py * p.W + px;
    const uint32_t rec_base = smem_u32(s_rec);
    uint32_t panel_base = smem_u32(s_panel) + warp * (32 * kPanelRow * 4);
    const uint32_t id_base = smem_u32(s_id);
    const uint32_t mask_base = smem_u32(s_mask) + (uint32_t)warp * (kGroupsB * 4);
    unsigned lt_mask = (1u << lane) - 1u;

    float T_final = 0, final_D = 0, final_D2 = 0;
    uint32_t last_contributor = 0, median_contributor = 0;
    float dpix0 = 0, dpix1 = 0, dpix2 = 0, dN0 = 0, dN1 = 0, dN2 = 0;
    float dL_ddepth = 0, dL_daccum = 0, dL_dreg = 0, dL_dmedian = 0;
    if (inside) {
        T_final = p.accum[pix]; final_D = p.accum[HW + pix]; final_D2 = p.accum[2 * HW + pix];
        last_contributor = p.n_contrib[pix]; median_contributor = p.n_contrib[HW + pix];
        const size_t GP = p.grad_plane;
        dpix0 = p.dL_dpix[pix]; dpix1 = p.dL_dpix[GP + pix]; dpix2 = p.dL_dpix[2 * GP + pix];
        dL_ddepth = p.dL_dothers[kChDepth * GP + pix];
        dL_daccum = p.dL_dothers[kChAlpha * GP + pix];
        dN0 = p.dL_dothers[(kChNormal + 0) * GP + pix];
        dN1 = p.dL_dothers[(kChNormal + 1) * GP + pix];
        dN2 = p.dL_dothers[(kChNormal + 2) * GP + pix];
        dL_dmedian = p.dL_dothers[kChMidDepth * GP + pix];
        dL_dreg = p.dL_dothers[kChDistortion * GP + pix];
    }
    // distortion terms folded with the pixel's cotangent once: dL/dw_j = D2r + m_j (m_j Ar - 2 D1r),
    // dL/dm_j = w_j (2 m_j Ar - 2 D1r)   (A = 1 - T_final, D1 = M1, D2 = M2 of the forward)
    const float Ar = (1.0f - T_final) * dL_dreg, Ar2 = Ar + Ar, nD1r2 = -2.0f * final_D * dL_dreg, D2r = final_D2 * dL_dreg;
    const uint32_t median_index = median_contributor - 1u;   // 0xFFFFFFFE when there is none
    const bool quirk = p.lowpass_quirk != 0;

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
    // into ONE accumulator S = sum_{j>i} w_j v_j (+ the background term):
    //     dL/dalpha_i = T_i v_i - S/(1-alpha_i).
    // S carries the background term as well: S = sum_{j>i} w_j v_j + T_final * bg . dL_dC
    float T = T_final;
    float S = T_final * ((__ldg(p.bg + 0) * dpix0 + __ldg(p.bg + 1) * dpix1) + __ldg(p.bg + 2) * dpix2);
    constexpr float kMScale = kFar / (kFar - kNear);
    constexpr float kDmScale = (kFar * kNear) / (kFar - kNear);

    for (int end = (int)cta_max; end > 0; end -= kBatchB) {
        const int n = min(kBatchB, end);
        const int start = end - n;
        if (end != (int)cta_max) __syncthreads();         // previous round's reads before this refill
        // ---- stage + classify (as in the forward) ----
#pragma unroll
        for (int k = 0; k < (kBatchB + 255) / 256; k++) {
            const int slot = k * 256 + tid;
            if (k * 256 + (warp << 5) >= n) break;                       // warp-uniform
            uint32_t m8 = 0;
            if (slot < n) {
                const uint32_t id = __ldg(p.point_list + range.x + start + slot);
                const float4* r = p.rec + (size_t)id * kRecQuads;
                const float4 bb = __ldg(r + 6), dg = __ldg(r + 7);
#ifdef SURFEL_STAGE_LDG
#pragma unroll
                for (int q = 0; q < kRecQuadsFwd; q++) s_rec[q * kBatchB + slot] = __ldg(r + q);
#else
#pragma unroll
                for (int q = 0; q < kRecQuadsFwd; q++) cp_async16(rec_base + (uint32_t)(q * kBatchB + slot) * 16u, r + q);
#endif
                s_id[slot] = id;
                m8 = classify_footprints(bb, dg, ox, oy);
            }
            uint32_t keep = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t b = __ballot_sync(0xffffffffu, (m8 >> w) & 1u);
                if (lane == w) keep = b;
            }
            if (lane < 8) s_mask[lane * kGroupsB + k * 8 + warp] = keep;
        }
        cp_async_wait_all();
        __syncthreads();
        if ((int)warp_max <= start) continue;   // nothing of this round reaches this warp's pixels

        const int top = min(n, (int)warp_max - start);            // slots [0, top) can contribute
        for (int g = (top - 1) >> 5; g >= 0; g--) {
            unsigned m = lds32u(mask_base + (uint32_t)g * 4u) & low_mask((uint32_t)(top - (g << 5)));
            uint32_t gb = rec_base + (uint32_t)(g << 5) * 16u, idb = id_base + (uint32_t)(g << 5) * 4u;
            // keep the shared-memory bases in registers (else re-derived from SR_CgaCtaId / SR_TID per hit)
            asm volatile("" : "+r"(gb), "+r"(idb), "+r"(panel_base), "+r"(lt_mask));
            const int own = (int)last_contributor - (start + (g << 5));      // bits below `own` are this pixel's
            const int med = (int)median_index - (start + (g << 5));
            while (m) {
                const uint32_t j = high_bit(m);                       // back to front
                m &= low_mask(j);
                const uint32_t ra = gb + j * 16u;
                const float4 q0 = lds128(ra), q1 = lds128(ra + kBatchB * 16), q2 = lds128(ra + 2 * kBatchB * 16);
                PairEval e;
                const bool active = eval_pair(pxf, pyf, q0, q1, q2, e) && (int)j < own;
                const unsigned am = __ballot_sync(0xffffffffu, active);
                if (am == 0u) continue;

                if (active) {
                    const float4 q3 = lds128(ra + 3 * kBatchB * 16), q4 = lds128(ra + 4 * kBatchB * 16);
                    const bool use3d = e.rho3d <= e.rho2d;
                    const float depth = use3d ? q4.w * e.inv_pz : q3.w;      // det T / p.z, or Tw.z in the low-pass branch
                    // one row per contributing lane (rows are compacted: ballot prefix)
                    const uint32_t ro = panel_base + (uint32_t)__popc(am & lt_mask) * (kPanelRow * 4);
                    // A.3's `depth < near` skip (rare: splats reaching through the near plane): the pair was not
                    // blended, so it leaves T and S alone and contributes a row of zeros
                    if (q2.w < 0.0f && depth < kNear) {            // flagged splats only (warp-uniform flag)
                        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        sts128(ro, z4); sts128(ro + 16, z4); sts128(ro + 32, z4); sts128(ro + 48, z4); sts128(ro + 64, z4);
                        sts64(ro + 80, 0.f, 0.f);
                    } else {
                        float ax = 0, ay = 0, az = 0, zd = 0, zx = 0, zy = 0, zz = 0, m2x = 0, m2y = 0;
                        const float G = e.G, alpha = e.alpha;
                        const float inv1ma = fast_rcp(1.0f - alpha);
                        T = T * inv1ma;
                        const float w = alpha * T;
                        const float inv_d = fast_rcp(depth);
                        const float m_d = fmaf(inv_d, -kMScale * kNear, kMScale);
                        const float dmd_dd = kDmScale * inv_d * inv_d;
                        float v = fmaf(m_d, fmaf(m_d, Ar, nD1r2), D2r) + dL_daccum;       // dL_dweight + dL_dA
                        v = fmaf(q4.x, dpix0, v); v = fmaf(q4.y, dpix1, v); v = fmaf(q4.z, dpix2, v);
                        v = fmaf(depth, dL_ddepth, v);
                        v = fmaf(q3.x, dN0, v); v = fmaf(q3.y, dN1, v); v = fmaf(q3.z, dN2, v);
                        const float dL_dalpha = fmaf(T, v, -(S * inv1ma));
                        S = fmaf(w, v, S);
                        float dL_dz = ((int)j == med) ? dL_dmedian : 0.0f;
                        dL_dz = fmaf(w * dmd_dd, fmaf(m_d, Ar2, nD1r2), dL_dz);
                        const float dL_dG = fabsf(q2.w) * dL_dalpha;
                        dL_dz = fmaf(w, dL_ddepth, dL_dz);
                        const float dop = G * dL_dalpha;
                        if (use3d) {
                            // G = exp(-0.5 |s|^2), s = p.xy / p.z, depth = det T / p.z
                            const float t = -G * dL_dG * e.inv_pz;
                            ax = t * e.sx; ay = t * e.sy;
                            zd = dL_dz * e.inv_pz;                                   // dL/d(det T)
                            az = -fmaf(ax, e.sx, fmaf(ay, e.sy, zd * depth));        // dL/dp.z
                        } else {
                            // upstream: dL_dmean2D += dL_dG * (-G * FilterInvSquare * d), d = c - pixel = -(dx, dy)
                            const float gg = G * kFilterInvSquare * dL_dG;
                            m2x = gg * e.dx; m2y = gg * e.dy;
                            zz = dL_dz;
                            // upstream "Propagate the gradients of depth" in this branch: dL_dTw += (s.x, s.y, 1) dL_dz
                            if (quirk) { zx = e.sx * dL_dz; zy = e.sy * dL_dz; }
                        }
                        sts128(ro, make_float4(ax, ay, az, e.dx * ax));
                        sts128(ro + 16, make_float4(e.dx * ay, e.dx * az, e.dy * ax, e.dy * ay));
                        sts128(ro + 32, make_float4(e.dy * az, zd, zx, zy));
                        sts128(ro + 48, make_float4(zz, m2x, m2y, dop));
                        sts128(ro + 64, make_float4(w * dN0, w * dN1, w * dN2, w * dpix0));
                        sts64(ro + 80, w * dpix1, w * dpix2);
                    }
                }
                const uint32_t id = lds32u(idb + j * 4u);          // requested before the column sums need it
                __syncwarp();
                {
                    // every lane sums a column (lanes 24..31 re-read column 23, padding, and drop the result): the
                    // loop runs converged, with a warp-uniform trip count
                    const float acc = column_sum(panel_base + (uint32_t)min(__popc(lt_mask), kPanelRow - 1) * 4u, __popc(am));
                    if (lane < kGradUsed && acc != 0.0f) {
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
    static bool attr_set[kMaxDevices] = {};
    const int slot = current_device_slot();
    if (slot < 0 || !attr_set[slot]) {
        SURFEL_CUDA_OK(cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
        if (slot >= 0) attr_set[slot] = true;
    }
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderBwd, stream);
    render_bwd_kernel<<<grid, 256, kBwdSmemBytes, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
