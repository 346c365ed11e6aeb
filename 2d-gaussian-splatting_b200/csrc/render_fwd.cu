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
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
#include "render_fwd_tile.inc"
    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.accum[pix] = T; p.accum[HW + pix] = M1; p.accum[2 * HW + pix] = M2;
        p.n_contrib[pix] = last_contributor & 0x7FFFFFFFu; p.n_contrib[HW + pix] = median_contributor;
        const size_t OP = p.out_plane;
        p.out_color[pix] = C0 + T * __ldg(p.bg + 0);
        p.out_color[OP + pix] = C1 + T * __ldg(p.bg + 1);
        p.out_color[2 * OP + pix] = C2 + T * __ldg(p.bg + 2);
        p.out_others[kChDepth * OP + pix] = D;
        p.out_others[kChAlpha * OP + pix] = 1.0f - T;
        p.out_others[(kChNormal + 0) * OP + pix] = N0;
        p.out_others[(kChNormal + 1) * OP + pix] = N1;
        p.out_others[(kChNormal + 2) * OP + pix] = N2;
        p.out_others[kChMidDepth * OP + pix] = median_depth;
        p.out_others[kChDistortion * OP + pix] = dist;
    }
}

// The ten output values of one pixel, in frame-plane order (3 colour planes, then the 7 planes of allmap).
struct PixelOut { float v[10]; };

// One tile for the pair kernel below: the same blend, the backward's per-pixel state written, outputs returned.
__device__ __forceinline__ void render_tile(const RenderParams& p, unsigned char* smem_raw, const int tx, const int ty, PixelOut& out) {
#include "render_fwd_tile.inc"
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
