// render_fwd.cu — per-tile front-to-back alpha blend, forward.
//
// Replaces upstream renderCUDA forward (algorithm SURVEY Appendix A.3): one CTA
// per 16x16 tile walks the tile's depth-sorted splat list and blends RGB, expected depth, alpha,
// view-space normal, median depth and the depth-distortion accumulator in one pass, saving
// final_T / M1 / M2 / n_contrib / median_contributor for the backward.
//
// B200 design (not upstream's):
//  * the tile's list is staged ONCE for all but crowded tiles (kBatch = 384 slots; the mean list at the
//    headline workload is 292), so the hit loop runs without a block barrier; longer lists take further
//    rounds;
//  * staging thread t gathers splat t's 128-byte record: quads 0-4 go straight to quad-planar shared memory
//    with cp.async (LDGSTS, no registers), the two culling quads into registers, where the thread classifies the splat against the
//    eight 8x4 warp footprints of the tile (screen AABB + diagonal extents of the region where alpha can
//    reach 1/255, render_common.cuh).  Ballots turn the classification into one 32-bit hit mask per
//    (warp, 32-slot group) in shared memory;
//  * each warp owns an 8x4 pixel footprint and evaluates only its hits: a miss cannot reach
//    alpha >= 1/255 on any of the warp's pixels, so skipping it is exact, and the contributor counter is
//    derived from the list position, so n_contrib / median contributor are unchanged.  Hit bits are
//    peeled front to back with FLO/BMSK, the ray-splat intersection costs 6 FMA (affine form, common.cuh);
//  * early termination per warp (all 32 pixels saturated), then per CTA.
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

#ifndef SURFEL_FWD_BATCH
#define SURFEL_FWD_BATCH 384
#endif
#ifndef SURFEL_FWD_BLOCKS
#define SURFEL_FWD_BLOCKS 5
#endif
constexpr int kBatch = SURFEL_FWD_BATCH;          // multiple of 32
constexpr int kGroups = kBatch / 32;
constexpr int kFwdSmemBytes = kRecQuadsFwd * kBatch * 16 + 8 * kGroups * 4;

__global__ void __launch_bounds__(256, SURFEL_FWD_BLOCKS) render_fwd_kernel(RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_rec = reinterpret_cast<float4*>(smem_raw);                                   // [quad][slot]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw + kRecQuadsFwd * kBatch * 16);  // [warp][group], bit-reversed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    int lx, ly;
    warp_pixel(warp, lane, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const float ox = (float)(tx * kBlockX), oy = (float)(ty * kBlockY);

    const uint2 range = p.ranges[ty * p.gx + tx];
    const int total = (int)(range.y - range.x);
    const uint32_t rec_base = smem_u32(s_rec);
    const uint32_t mask_base = smem_u32(s_mask) + (uint32_t)warp * (kGroups * 4);
    constexpr float kMScale = kFar / (kFar - kNear);

    float T = 1.0f, C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0, D = 0, M1 = 0, M2 = 0, dist = 0;
    float median_depth = 0;
    uint32_t last_contributor = inside ? 0u : 0x80000000u;   // top bit: this pixel is finished
    uint32_t median_contributor = 0xFFFFFFFFu;
    bool warp_done = __all_sync(0xffffffffu, !inside);

    for (int base = 0; base < total; base += kBatch) {
        // CTA-wide early out (also orders the previous round's smem reads before this refill)
        if (base > 0 && !__syncthreads_or(!warp_done)) break;

        const int n = min(kBatch, total - base);
        // ---- stage + classify: thread t takes slots t, t + 256, ... (whole warps stay together) ----
#pragma unroll
        for (int k = 0; k < (kBatch + 255) / 256; k++) {
            const int slot = k * 256 + tid;
            if (k * 256 + (warp << 5) >= n) break;                       // warp-uniform
            uint32_t m8 = 0;
            if (slot < n) {
                const uint32_t id = __ldg(p.point_list + range.x + base + slot);
                const float4* r = p.rec + (size_t)id * kRecQuads;
                const float4 bb = __ldg(r + 6), dg = __ldg(r + 7);
#ifdef SURFEL_STAGE_LDG
#pragma unroll
                for (int q = 0; q < kRecQuadsFwd; q++) s_rec[q * kBatch + slot] = __ldg(r + q);
#else
#pragma unroll
                for (int q = 0; q < kRecQuadsFwd; q++) cp_async16(rec_base + (uint32_t)(q * kBatch + slot) * 16u, r + q);
#endif
                m8 = classify_footprints(bb, dg, ox, oy);
            }
            uint32_t keep = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t b = __ballot_sync(0xffffffffu, (m8 >> w) & 1u);
                if (lane == w) keep = b;
            }
            // bit 31 = first slot of the group: taking the highest set bit first walks the hits front to back
            if (lane < 8) s_mask[lane * kGroups + k * 8 + warp] = __brev(keep);
        }
        cp_async_wait_all();
        __syncthreads();

        if (!warp_done) {
            const int ngroups = (n + 31) >> 5;
            for (int g = 0; g < ngroups; g++) {
                unsigned m = lds32u(mask_base + (uint32_t)g * 4u);
                // The hit loop holds no warp-synchronous operation, so finished pixels simply skip it
                // and a pixel that saturates leaves it early.  "Finished" is the top bit of
                // last_contributor (lists are far shorter than 2^31).
                if ((int)last_contributor >= 0) {
                    const uint32_t g31 = rec_base + (uint32_t)(g * 32 + 31) * 16u;
                    const uint32_t k32 = (uint32_t)(base + g * 32 + 32);
                    while (m) {
                        const uint32_t hb = high_bit(m);               // slot g*32 + 31 - hb
                        m &= low_mask(hb);
                        const uint32_t ra = g31 - hb * 16u;
                        const float4 q0 = lds128(ra), q1 = lds128(ra + kBatch * 16), q2 = lds128(ra + 2 * kBatch * 16);
                        PairEval e;
                        if (!eval_pair(pxf, pyf, q0, q1, q2, e)) continue;
                        const float4 q3 = lds128(ra + 3 * kBatch * 16), q4 = lds128(ra + 4 * kBatch * 16);
                        // ray-splat depth = w of the intersection = det T / p.z; low-pass branch: Tw.z
                        const float depth = (e.rho3d <= e.rho2d) ? q4.w * e.inv_pz : q3.w;
                        if (q2.w < 0.0f && depth < kNear) continue;    // flagged splats only (warp-uniform flag)
                        const float test_T = T * (1.0f - e.alpha);
                        if (test_T < kTMin) { last_contributor |= 0x80000000u; break; }
                        const uint32_t contributor = k32 - hb;   // 1-based list position
                        const float w = e.alpha * T;
                        const float A = 1.0f - T;
                        const float mm = fmaf(fast_rcp(depth), -kMScale * kNear, kMScale);
                        dist = fmaf(fmaf(mm, fmaf(mm, A, -(M1 + M1)), M2), w, dist);     // (mm^2 A + M2 - 2 mm M1) w
                        D += depth * w;
                        M1 += mm * w;
                        M2 += mm * mm * w;
                        if (T > 0.5f) { median_depth = depth; median_contributor = contributor; }
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

    const float c0 = C0 + T * __ldg(p.bg + 0), c1 = C1 + T * __ldg(p.bg + 1), c2 = C2 + T * __ldg(p.bg + 2);
    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.accum[pix] = T; p.accum[HW + pix] = M1; p.accum[2 * HW + pix] = M2;
        p.n_contrib[pix] = last_contributor & 0x7FFFFFFFu; p.n_contrib[HW + pix] = median_contributor;
        if (p.rep_count == 0) {
            const size_t OP = p.out_plane;
            p.out_color[pix] = c0;
            p.out_color[OP + pix] = c1;
            p.out_color[2 * OP + pix] = c2;
            p.out_others[kChDepth * OP + pix] = D;
            p.out_others[kChAlpha * OP + pix] = 1.0f - T;
            p.out_others[(kChNormal + 0) * OP + pix] = N0;
            p.out_others[(kChNormal + 1) * OP + pix] = N1;
            p.out_others[(kChNormal + 2) * OP + pix] = N2;
            p.out_others[kChMidDepth * OP + pix] = median_depth;
            p.out_others[kChDistortion * OP + pix] = dist;
        }
    }
    if (p.rep_count != 0) {
        // Tile-band exchange fused into the producer: the band's pixels go straight to every replica of the
        // frame — peer GPUs' memory over NVLink, or one NVSwitch multicast address that the switch fans out
        // (a plain st.global to a multicast mapping IS multimem.st: same SASS) — fire-and-forget stores that
        // overlap with the blending of the CTAs still running.  No all-gather follows; the caller runs a
        // cross-GPU barrier before reading rows of other bands.
        // The tile's ten planes are first transposed through shared memory (the record slab is dead by now) so
        // that a warp's store covers two full 64-byte tile rows instead of four 32-byte footprint rows: remote
        // writes are limited by the number of requests the receiving GPU can take, not by bytes (measured:
        // ~225 GB/s inbound with 32-byte segments whatever the number of senders).
        __syncthreads();                                   // every warp has left the hit loop
        float* s_out = reinterpret_cast<float*>(smem_raw); // [plane][16][16]
        const int o = ly * kBlockX + lx;
        s_out[0 * 256 + o] = c0; s_out[1 * 256 + o] = c1; s_out[2 * 256 + o] = c2;
        s_out[(3 + kChDepth) * 256 + o] = D; s_out[(3 + kChAlpha) * 256 + o] = 1.0f - T;
        s_out[(3 + kChNormal + 0) * 256 + o] = N0; s_out[(3 + kChNormal + 1) * 256 + o] = N1;
        s_out[(3 + kChNormal + 2) * 256 + o] = N2;
        s_out[(3 + kChMidDepth) * 256 + o] = median_depth; s_out[(3 + kChDistortion) * 256 + o] = dist;
        __syncthreads();
        const int rx = tid & 15, ry = tid >> 4;            // row-major over the tile: a warp = two 16-pixel rows
        const int gx2 = tx * kBlockX + rx, gy2 = ty * kBlockY + ry;
        if (gx2 < p.W && gy2 < p.H) {
            const size_t OP = p.out_plane;
            const size_t pix2 = (size_t)gy2 * p.W + gx2;
            float v[10];
#pragma unroll
            for (int c = 0; c < 10; c++) v[c] = s_out[c * 256 + tid];
            for (int r = 0; r < p.rep_count; r++) {
                float* b = reinterpret_cast<float*>(p.rep_base[r]) + pix2;
#pragma unroll
                for (int c = 0; c < 10; c++) b[c * OP] = v[c];
            }
        }
    }
}

int launch_render_fwd(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    static bool attr_set[kMaxDevices] = {};
    const int slot = current_device_slot();
    if (slot < 0 || !attr_set[slot]) {
        SURFEL_CUDA_OK(cudaFuncSetAttribute(render_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes));
        if (slot >= 0) attr_set[slot] = true;
    }
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderFwd, stream);
    render_fwd_kernel<<<grid, 256, kFwdSmemBytes, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
