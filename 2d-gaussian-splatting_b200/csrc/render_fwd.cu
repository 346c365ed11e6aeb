// render_fwd.cu — per-tile front-to-back alpha blend, forward.
//
// Replaces upstream renderCUDA forward (algorithm SURVEY Appendix A.3): one CTA
// per 16x16 tile walks the tile's depth-sorted splat list and blends RGB, expected depth, alpha,
// view-space normal, median depth and the depth-distortion accumulator in one pass, saving
// final_T / M1 / M2 / n_contrib / median_contributor for the backward.
//
// B200 design (not upstream's): the list is staged 256 splats at a time as 96-byte records
// (6 x LDG.128 per splat -> conflict-free quad-planar shared memory).  Each warp owns an 8x4 pixel
// footprint; its 32 lanes test 32 staged splats at once against the footprint using the splat's
// conservative screen bbox (record quad 5), ballot the hits, and only the hit splats are evaluated
// (warp-wide compaction).  A splat whose bbox misses the footprint cannot reach alpha >= 1/255 on
// any of the warp's pixels, so skipping it is exact, not approximate; the contributor counter is
// derived from the list position, so bookkeeping (n_contrib, median contributor) is unchanged.
// Early termination is per warp (all 32 pixels done), then per CTA.
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

#ifndef SURFEL_FWD_BATCH
#define SURFEL_FWD_BATCH 256
#endif
constexpr int kBatch = SURFEL_FWD_BATCH;

template <bool kSlab>   // kSlab: also lay the staged records out in sorted order (TMA-backward variant only)
__global__ void __launch_bounds__(256, 5) render_fwd_kernel(RenderParams p) {
    __shared__ float4 s_rec[kRecQuads * kBatch];     // [quad][slot]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    int lx, ly;
    warp_pixel(warp, lane, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    // warp footprint (pixel-centre coordinates)
    const float fx0 = (float)(tx * kBlockX + ((warp & 1) << 3)), fx1 = fx0 + 7.0f;
    const float fy0 = (float)(ty * kBlockY + ((warp >> 1) << 2)), fy1 = fy0 + 3.0f;

    const uint2 range = p.ranges[ty * p.gx + tx];
    const int total = (int)(range.y - range.x);
    const uint32_t rec_base = smem_u32(s_rec);
    constexpr float kMScale = kFar / (kFar - kNear);

    float T = 1.0f, C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0, D = 0, M1 = 0, M2 = 0, dist = 0;
    float median_depth = 0;
    uint32_t last_contributor = inside ? 0u : 0x80000000u;   // top bit: this pixel is finished
    uint32_t median_contributor = 0xFFFFFFFFu;
    bool warp_done = __all_sync(0xffffffffu, !inside);

    for (int base = 0; base < total; base += kBatch) {
        // CTA-wide early out (also orders the previous batch's smem reads before this refill)
        if (!__syncthreads_or(!warp_done)) break;

        const int n = min(kBatch, total - base);
        if (tid < n) {
            const uint32_t id = p.point_list[range.x + base + tid];
            const float4* r = p.rec + (size_t)id * kRecQuads;
#pragma unroll
            for (int q = 0; q < kRecQuads; q++) s_rec[q * kBatch + tid] = __ldg(r + q);
            if (kSlab) {
                float4* dst = p.slab + (size_t)(range.x + base + tid) * kRecQuads;
#pragma unroll
                for (int q = 0; q < kRecQuads; q++) dst[q] = s_rec[q * kBatch + tid];
            }
        }
        __syncthreads();

        if (!warp_done) {
            for (int c = 0; c < n; c += 32) {
                const int slot = c + lane;
                bool hit = false;
                if (slot < n) {
                    const float4 bb = lds128(rec_base + (5 * kBatch + slot) * 16);
                    hit = bb.x <= fx1 && bb.z >= fx0 && bb.y <= fy1 && bb.w >= fy0;
                }
                // bit 31 = slot c: taking the highest set bit first walks the hits front to back
                unsigned m = __brev(__ballot_sync(0xffffffffu, hit));
                // The hit loop holds no warp-synchronous operation, so finished pixels simply skip it
                // and a pixel that saturates leaves it early.  "Finished" is the top bit of
                // last_contributor (lists are far shorter than 2^31).
                if ((int)last_contributor >= 0) {
                    const uint32_t g31 = rec_base + (uint32_t)(c + 31) * 16u;
                    const uint32_t k32 = (uint32_t)(base + c + 32);
                    while (m) {
                        const uint32_t hb = high_bit(m);               // slot c + 31 - hb
                        m &= low_mask(hb);
                        const uint32_t ra = g31 - hb * 16u;
                        const float4 q0 = lds128(ra), q1 = lds128(ra + kBatch * 16), q2 = lds128(ra + 2 * kBatch * 16);
                        PairEval e;
                        if (!eval_pair(pxf, pyf, q0, q1, q2, e)) continue;
                        const float test_T = T * (1.0f - e.alpha);
                        if (test_T < kTMin) { last_contributor |= 0x80000000u; break; }
                        const uint32_t contributor = k32 - hb;   // 1-based list position
                        const float4 q3 = lds128(ra + 3 * kBatch * 16), q4 = lds128(ra + 4 * kBatch * 16);
                        const float w = e.alpha * T;
                        const float A = 1.0f - T;
                        const float mm = kMScale * (1.0f - kNear * fast_rcp(e.depth));
                        dist += (mm * mm * A + M2 - 2.0f * mm * M1) * w;
                        D += e.depth * w;
                        M1 += mm * w;
                        M2 += mm * mm * w;
                        if (T > 0.5f) { median_depth = e.depth; median_contributor = contributor; }
                        N0 += q3.x * w; N1 += q3.y * w; N2 += q3.z * w;
                        C0 += q4.x * w; C1 += q4.y * w; C2 += q4.z * w;
                        T = test_T;
                        last_contributor = contributor;
                    }
                }
                if (__all_sync(0xffffffffu, (int)last_contributor < 0)) { warp_done = true; break; }
            }
        }
    }

    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.accum[pix] = T; p.accum[HW + pix] = M1; p.accum[2 * HW + pix] = M2;
        p.n_contrib[pix] = last_contributor & 0x7FFFFFFFu; p.n_contrib[HW + pix] = median_contributor;
        p.out_color[pix] = C0 + T * __ldg(p.bg + 0);
        p.out_color[HW + pix] = C1 + T * __ldg(p.bg + 1);
        p.out_color[2 * HW + pix] = C2 + T * __ldg(p.bg + 2);
        p.out_others[kChDepth * HW + pix] = D;
        p.out_others[kChAlpha * HW + pix] = 1.0f - T;
        p.out_others[(kChNormal + 0) * HW + pix] = N0;
        p.out_others[(kChNormal + 1) * HW + pix] = N1;
        p.out_others[(kChNormal + 2) * HW + pix] = N2;
        p.out_others[kChMidDepth * HW + pix] = median_depth;
        p.out_others[kChDistortion * HW + pix] = dist;
    }
}

int launch_render_fwd(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderFwd, stream);
    if (p.slab) render_fwd_kernel<true><<<grid, 256, 0, stream>>>(p);
    else        render_fwd_kernel<false><<<grid, 256, 0, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
