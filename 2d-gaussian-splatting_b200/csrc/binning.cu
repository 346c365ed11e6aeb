// binning.cu — (tile | depth) key emission and tile range detection.
//
// Replaces upstream duplicateWithKeys and identifyTileRanges (algorithm
// SURVEY Appendix A.2).  Keys are (tile_id << 32) | float_bits(view depth); values are splat ids;
// the emission order (splat index ascending, then tile rows, then columns) is kept exactly because
// the stable sort's tie order depends on it.
//
// B200 notes: upstream runs one thread per splat with a serial loop over its tiles (load imbalance,
// scattered 12-byte writes).  Here a block owns 256 consecutive splats and its threads walk the
// block's FLATTENED instance range, so every store is fully coalesced and a splat covering
// thousands of tiles is spread over the whole block.
#include "common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

constexpr int kDupBlock = 256;

__global__ void __launch_bounds__(kDupBlock)
duplicate_with_keys_kernel(int P, int gx, int gy, int row0, int row1, const float4* __restrict__ tmat,
                           const int* __restrict__ radii, const uint32_t* __restrict__ offsets,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint32_t s_end[kDupBlock];
    __shared__ int s_x0[kDupBlock], s_y0[kDupBlock], s_w[kDupBlock];
    __shared__ uint32_t s_depth[kDupBlock];
    const int tid = threadIdx.x;
    const int first = blockIdx.x * kDupBlock;
    const int idx = first + tid;
    const uint32_t base = first == 0 ? 0u : offsets[first - 1];
    const int last = min(P, first + kDupBlock) - 1;
    s_end[tid] = offsets[min(idx, last)];
    int x0 = 0, y0 = 0, w = 1;
    uint32_t dbits = 0;
    if (idx < P) {
        const int r = radii[idx];
        if (r > 0) {
            const float4 t2 = tmat[(size_t)idx * kTmQuads + 2];      // (Tw.z, xy.x, xy.y, view depth)
            int x1, y1;
            get_rect(t2.y, t2.z, r, gx, gy, row0, row1, x0, y0, x1, y1);
            w = max(1, x1 - x0);
            dbits = __float_as_uint(t2.w);
        }
    }
    s_x0[tid] = x0; s_y0[tid] = y0; s_w[tid] = w; s_depth[tid] = dbits;
    __syncthreads();
    const uint32_t total = s_end[kDupBlock - 1] - base;
    for (uint32_t i = tid; i < total; i += kDupBlock) {
        const uint32_t g = base + i;
        // first splat s in the block with s_end[s] > g
        int lo = 0, hi = kDupBlock - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_end[mid] > g) hi = mid; else lo = mid + 1;
        }
        const uint32_t start = lo == 0 ? base : s_end[lo - 1];
        const uint32_t j = g - start;
        const uint32_t ww = (uint32_t)s_w[lo];
        const uint32_t ry = j / ww, rx = j - ry * ww;
        const uint32_t tile = (uint32_t)(s_y0[lo] + (int)ry) * (uint32_t)gx + (uint32_t)(s_x0[lo] + (int)rx);
        keys[g] = ((uint64_t)tile << 32) | (uint64_t)s_depth[lo];
        vals[g] = (uint32_t)(first + lo);
    }
}

__global__ void identify_tile_ranges_kernel(size_t R, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[tile].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (prev != tile) { ranges[prev].y = (uint32_t)i; ranges[tile].x = (uint32_t)i; }
    }
    if (i == R - 1) ranges[tile].y = (uint32_t)R;
}

int launch_duplicate_with_keys(int P, int gx, int gy, int row0, int row1, const float4* tmat,
                               const int* radii, const uint32_t* offsets, uint64_t* keys,
                               uint32_t* vals, cudaStream_t stream) {
    if (P <= 0) return 0;
    LaunchScope scope(kStDuplicate, stream);
    duplicate_with_keys_kernel<<<(P + kDupBlock - 1) / kDupBlock, kDupBlock, 0, stream>>>(
        P, gx, gy, row0, row1, tmat, radii, offsets, keys, vals);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_identify_tile_ranges(size_t R, int tiles, const uint64_t* keys_sorted, uint2* ranges,
                                cudaStream_t stream) {
    SURFEL_CUDA_OK(cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), stream));
    if (R == 0) return 0;
    LaunchScope scope(kStRanges, stream);
    identify_tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(R, keys_sorted, ranges);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
