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

// The ten output values of one pixel, in frame-plane order (3 colour planes, then the 7 planes of allmap).
struct PixelOut { float v[10]; };

// One 16x16 tile, front to back.  Writes the backward's per-pixel state (accum, n_contrib) and returns the
// pixel's outputs; where they are stored is the caller's business.
__device__ __forceinline__ void render_tile(const RenderParams& p, unsigned char* smem_raw, const int tx, const int ty, PixelOut& out) {
    float4* s_rec = reinterpret_cast<float4*>(smem_raw);                                   // [quad][slot]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw + kRecQuadsFwd * kBatch * 16);  // [warp][group], bit-reversed

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.accum[pix] = T; p.accum[HW + pix] = M1; p.accum[2 * HW + pix] = M2;
        p.n_contrib[pix] = last_contributor & 0x7FFFFFFFu; p.n_contrib[HW + pix] = median_contributor;
    }
    out.v[0] = C0 + T * __ldg(p.bg + 0); out.v[1] = C1 + T * __ldg(p.bg + 1); out.v[2] = C2 + T * __ldg(p.bg + 2);
    out.v[3 + kChDepth] = D; out.v[3 + kChAlpha] = 1.0f - T;
    out.v[3 + kChNormal + 0] = N0; out.v[3 + kChNormal + 1] = N1; out.v[3 + kChNormal + 2] = N2;
    out.v[3 + kChMidDepth] = median_depth; out.v[3 + kChDistortion] = dist;
}

__global__ void __launch_bounds__(256, SURFEL_FWD_BLOCKS) render_fwd_kernel(RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    PixelOut o;
    render_tile(p, smem_raw, tx, ty, o);
    int lx, ly;
    warp_pixel(threadIdx.x >> 5, threadIdx.x & 31, lx, ly);
    const int px = tx * kBlockX + lx, py = ty * kBlockY + ly;
    if (px < p.W && py < p.H) {
        const size_t OP = p.out_plane;
        const size_t pix = (size_t)py * p.W + px;
        p.out_color[pix] = o.v[0];
        p.out_color[OP + pix] = o.v[1];
        p.out_color[2 * OP + pix] = o.v[2];
#pragma unroll
        for (int c = 0; c < 7; c++) p.out_others[c * OP + pix] = o.v[3 + c];
    }
}

// Tile-band exchange fused into the producer (surfel_settings.out_replica_base): the band's pixels go straight to
// every replica of the frame — peer GPUs' memory over NVLink, or one NVSwitch multicast address that the switch
// fans out (a plain st.global to a multicast mapping IS multimem.st: same SASS) — fire-and-forget stores that
// overlap with the blending of the CTAs still running.  No all-gather follows; the caller runs a cross-GPU
// barrier before reading rows of other bands.
// Remote writes are limited by the number of REQUESTS the receiving GPU can take, not by bytes (measured at
// N = 2 / 8: ~225 GB/s inbound with the 32-byte rows of a warp footprint whatever the number of senders, twice
// that with 64-byte tile rows).  So one CTA renders TWO horizontally adjacent tiles, parks their outputs in
// shared memory (the first tile's beside the record slab, the second's in the slab once it is dead), and every
// warp store then covers one full 128-byte row of the 32-pixel strip.
constexpr int kPairOutBytes = 10 * 256 * 4;                       // one tile's ten planes
constexpr int kFwdPairSmemBytes = kFwdSmemBytes + kPairOutBytes;

__global__ void __launch_bounds__(256, SURFEL_FWD_BLOCKS) render_fwd_pair_kernel(RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_a = reinterpret_cast<float*>(smem_raw + kFwdSmemBytes);     // first tile  [plane][16][16]
    float* s_b = reinterpret_cast<float*>(smem_raw);                      // second tile, over the dead slab
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * 2, ty = blockIdx.y + p.row0;
    const bool two = tx0 + 1 < p.gx;
    int lx, ly;
    warp_pixel(tid >> 5, tid & 31, lx, ly);
    const int o16 = ly * kBlockX + lx;
    {
        PixelOut o;
        render_tile(p, smem_raw, tx0, ty, o);
#pragma unroll
        for (int c = 0; c < 10; c++) s_a[c * 256 + o16] = o.v[c];
    }
    __syncthreads();                                   // every warp has left the first tile's slab
    if (two) {
        PixelOut o;
        render_tile(p, smem_raw, tx0 + 1, ty, o);
        __syncthreads();                               // ... and the second's
#pragma unroll
        for (int c = 0; c < 10; c++) s_b[c * 256 + o16] = o.v[c];
    }
    __syncthreads();
    // warp w writes rows w and w + 8 of the strip: lane = pixel x of the 32-pixel row
    const int lane = tid & 31, warp = tid >> 5;
    const int gx2 = tx0 * kBlockX + lane;
    const float* src = lane < 16 ? s_a : s_b;
    const int sx = lane & 15;
    const size_t OP = p.out_plane;
    if (gx2 < p.W && (lane < 16 || two)) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int ry = warp + 8 * h, gy2 = ty * kBlockY + ry;
            if (gy2 >= p.H) continue;
            const size_t pix2 = (size_t)gy2 * p.W + gx2;
            float v[10];
#pragma unroll
            for (int c = 0; c < 10; c++) v[c] = src[c * 256 + ry * kBlockX + sx];
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
        SURFEL_CUDA_OK(cudaFuncSetAttribute(render_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdPairSmemBytes));
        if (slot >= 0) attr_set[slot] = true;
    }
    LaunchScope scope(kStRenderFwd, stream);
    if (p.rep_count > 0) render_fwd_pair_kernel<<<dim3((p.gx + 1) / 2, rows), 256, kFwdPairSmemBytes, stream>>>(p);
    else                 render_fwd_kernel<<<dim3(p.gx, rows), 256, kFwdSmemBytes, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
