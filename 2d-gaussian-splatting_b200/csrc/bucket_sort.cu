// bucket_sort.cu — tile-bucketed binning: the production replacement for
// duplicateWithKeys + device-wide radix sort + identifyTileRanges.
//
// The required result is the instance list sorted by (tile, depth bits, splat index) plus the
// per-tile [start,end) ranges.  A device-wide LSD radix sort reaches it with 6-7 passes over
// (u64,u32) pairs; on B200 at R ~ 2-3 M those passes are latency-bound (~38 us each, profiles/r1).
// The structure of the key allows a two-level scheme with ONE scatter and ONE local sort:
//   1. tile_count   : walk every block's flattened instance range, RED.add on the tile counters;
//   2. tile_scan    : exclusive scan of the counters -> ranges (identifyTileRanges for free),
//                     list of tiles too large for the fast sort;
//   3. tile_scatter : same walk; claim a slot in the tile's bucket with an atomic and store
//                     (depth_bits << 32 | splat_idx).  Arrival order inside a bucket is arbitrary;
//   4. tile_sort    : one CTA per tile sorts its bucket in shared memory (bitonic network on the
//                     64-bit (depth, idx) key — a total order, so the result is deterministic and
//                     equals the stable radix sort's: depth ascending, ties by splat index) and
//                     writes the point list (and, for tests, the full (tile|depth) keys);
//      tiles above kSmallMax entries use a 1024-thread CTA with up to 128 KB of shared memory, and
//      beyond that a global-memory bitonic sort (pathological inputs only; correct, slow).
// Traffic: 8 B written + 8 B read + 4 B written per instance instead of ~150 B.
#include <algorithm>
#include "common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

constexpr int kWalkBlock = 256;
constexpr int kWarpMax = 512;            // entries sorted by ONE WARP entirely in registers (16 per lane)
constexpr int kSmallMax = 2048;          // entries sorted by the 128-thread fast path (16 KB smem)
constexpr int kLargeMax = 16384;         // entries sorted in 128 KB of dynamic smem by 1024 threads

// Shared helper: block of 256 splats -> flattened instance walk.  F(tile, depth_bits, splat_idx).
template <typename F>
__device__ __forceinline__ void walk_instances(int P, int gx, int gy, int row0, int row1,
                                               const float4* __restrict__ tmat, const int* __restrict__ radii,
                                               const uint32_t* __restrict__ offsets, F f) {
    __shared__ uint32_t s_end[kWalkBlock];
    __shared__ int s_x0[kWalkBlock], s_y0[kWalkBlock], s_w[kWalkBlock];
    __shared__ uint32_t s_depth[kWalkBlock];
    const int tid = threadIdx.x;
    const int first = blockIdx.x * kWalkBlock;
    const int idx = first + tid;
    const uint32_t base = first == 0 ? 0u : offsets[first - 1];
    const int last = min(P, first + kWalkBlock) - 1;
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
    const uint32_t total = s_end[kWalkBlock - 1] - base;
    for (uint32_t i = tid; i < total; i += kWalkBlock) {
        const uint32_t g = base + i;
        int lo = 0, hi = kWalkBlock - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_end[mid] > g) hi = mid; else lo = mid + 1;
        }
        const uint32_t start = lo == 0 ? base : s_end[lo - 1];
        const uint32_t j = g - start;
        const uint32_t ww = (uint32_t)s_w[lo];
        const uint32_t ry = j / ww, rx = j - ry * ww;
        const uint32_t tile = (uint32_t)(s_y0[lo] + (int)ry) * (uint32_t)gx + (uint32_t)(s_x0[lo] + (int)rx);
        f(tile, s_depth[lo], (uint32_t)(first + lo));
    }
}

__global__ void __launch_bounds__(kWalkBlock)
tile_count_kernel(int P, int gx, int gy, int row0, int row1, const float4* __restrict__ tmat,
                  const int* __restrict__ radii, const uint32_t* __restrict__ offsets,
                  uint32_t* __restrict__ tile_count) {
    walk_instances(P, gx, gy, row0, row1, tmat, radii, offsets,
                   [&](uint32_t tile, uint32_t, uint32_t) { atomicAdd(tile_count + tile, 1u); });
}

// One block: exclusive scan of tile_count -> ranges; zero the fill cursors; collect big tiles.
// 4096 tiles per round (two rounds for a 1920x1080 frame, 32 for 7680x4320): every thread scans four CONSECUTIVE
// tiles, warp scans for both levels, the running carry lives in a register of every thread; the per-tile
// prefixes go through shared memory so that the global loads and stores of the emit phase are coalesced
// (thread-consecutive emission wrote 32 different sectors per warp store: 0.14-0.18 ms at 130 K tiles).
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int tiles, uint32_t cap, const uint32_t* __restrict__ tile_count, uint2* __restrict__ ranges,
                 uint32_t* __restrict__ tile_fill, uint32_t* __restrict__ big_list,
                 uint32_t* __restrict__ big_count, uint32_t* __restrict__ mid_list,
                 uint32_t* __restrict__ mid_count) {
    __shared__ uint32_t s_pre[4096];
    __shared__ uint32_t s_warp[32], s_warp_ex[32];
    __shared__ uint32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { *big_count = 0; *mid_count = 0; }
    uint32_t carry = 0;
    for (int base = 0; base < tiles; base += 4096) {
        const int t0 = base + tid * 4;
        uint32_t c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) c[k] = t0 + k < tiles ? tile_count[t0 + k] : 0u;
        const uint32_t sum = c[0] + c[1] + c[2] + c[3];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();                                   // also: the previous round's readers of s_pre are done
        if (warp == 0) {
            const uint32_t v = s_warp[lane];
            uint32_t vi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) vi += u;
            }
            s_warp_ex[lane] = vi - v;
            if (lane == 31) s_total = vi;
        }
        __syncthreads();
        uint32_t pre = carry + s_warp_ex[warp] + incl - sum;       // exclusive prefix of this thread's first tile
        carry += s_total;
#pragma unroll
        for (int k = 0; k < 4; k++) { s_pre[tid * 4 + k] = pre; pre += c[k]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int j = tid + k * 1024, t = base + j;
            if (t < tiles) {
                // `cap` = number of instance slots the caller allocated.  When the forward is launched
                // speculatively with a capacity guess (no host sync on R) and the guess was too small,
                // everything is clamped so that no kernel touches memory past the buffers; the host then
                // sees R > cap and re-runs with exact sizes.
                const uint32_t p0 = s_pre[j], cnt = tile_count[t];
                const uint32_t start = min(p0, cap), end = min(p0 + cnt, cap);
                const uint32_t cc = end - start;
                ranges[t] = cc ? make_uint2(start, end) : make_uint2(0u, 0u);
                tile_fill[t] = p0;
                if (cc > (uint32_t)kSmallMax) big_list[atomicAdd(big_count, 1u)] = (uint32_t)t;
                else if (cc > (uint32_t)kWarpMax) mid_list[atomicAdd(mid_count, 1u)] = (uint32_t)t;
            }
        }
    }
}

__global__ void __launch_bounds__(kWalkBlock)
tile_scatter_kernel(int P, int gx, int gy, int row0, int row1, const float4* __restrict__ tmat,
                    const int* __restrict__ radii, const uint32_t* __restrict__ offsets,
                    uint32_t* __restrict__ tile_fill, unsigned long long* __restrict__ pairs, uint32_t cap) {
    walk_instances(P, gx, gy, row0, row1, tmat, radii, offsets,
                   [&](uint32_t tile, uint32_t dbits, uint32_t idx) {
                       const uint32_t slot = atomicAdd(tile_fill + tile, 1u);
                       if (slot < cap) pairs[slot] = ((unsigned long long)dbits << 32) | idx;
                   });
}

// Bitonic sorting network in its all-ascending form: each merge of block size k starts with a
// "flip" step (i <-> mirror of i inside the block) followed by the usual half-cleaners j = k/4..1.
// Every compare-exchange moves the smaller key to the lower index, so elements beyond n can be
// treated as +inf without ever being touched: m is just the next power of two >= n.
// pair_of(): the t-th pair (i < l) of a step.
template <typename I>
__device__ __forceinline__ void flip_pair(I t, I k, int lg_half, I& i, I& l) {
    const I half = k >> 1, blk = t >> lg_half, o = t & (half - 1);
    i = blk * k + o;
    l = blk * k + (k - 1 - o);
}
template <typename I>
__device__ __forceinline__ void half_pair(I t, I j, I& i, I& l) {
    i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    l = i | j;
}

template <int T, typename I, typename Ptr>
__device__ __forceinline__ void bitonic_ascending(Ptr s, I n, I m, int tid) {
    int lg_half = 0;
    for (I k = 2; k <= m; k <<= 1, lg_half++) {
        for (I t = tid; t < (m >> 1); t += T) {
            I i, l;
            flip_pair<I>(t, k, lg_half, i, l);
            if (l < n) {
                const unsigned long long a = s[i], b = s[l];
                if (a > b) { s[i] = b; s[l] = a; }
            }
        }
        __syncthreads();
        for (I j = k >> 2; j > 0; j >>= 1) {
            for (I t = tid; t < (m >> 1); t += T) {
                I i, l;
                half_pair<I>(t, j, i, l);
                if (l < n) {
                    const unsigned long long a = s[i], b = s[l];
                    if (a > b) { s[i] = b; s[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    return n <= 32 ? 32 : 1 << (32 - __clz(n - 1));
}

template <int T>
__device__ __forceinline__ void sort_tile_in_smem(unsigned long long* s, uint32_t tile, uint2 rg,
                                                  const unsigned long long* __restrict__ pairs,
                                                  uint32_t* __restrict__ point_list,
                                                  unsigned long long* __restrict__ keys_sorted, int tid) {
    const int n = (int)(rg.y - rg.x);
    const int m = next_pow2(n);
    for (int i = tid; i < n; i += T) s[i] = pairs[rg.x + i];
    __syncthreads();
    bitonic_ascending<T, int>(s, n, m, tid);
    for (int i = tid; i < n; i += T) {
        const unsigned long long v = s[i];
        point_list[rg.x + i] = (uint32_t)v;
        if (keys_sorted) keys_sorted[rg.x + i] = ((unsigned long long)tile << 32) | (v >> 32);
    }
}

// ---- fast path for the common case (n <= 2048): 128 threads, E keys per thread in REGISTERS ----
// Classic bitonic network on the blocked layout e = tid*E + r: strides below E are compare-exchanges
// between registers of one thread, strides below 32*E are warp shuffles, only the few largest strides
// go through shared memory.  ~3x fewer instructions than running every stage through shared memory.
constexpr unsigned long long kPad = ~0ull;

__device__ __forceinline__ void ce_regs(unsigned long long& a, unsigned long long& b, bool asc) {
    const bool sw = (a > b) == asc;
    const unsigned long long lo = sw ? b : a, hi = sw ? a : b;
    a = lo; b = hi;
}

template <int E>
__device__ __forceinline__ void block_bitonic_regs(unsigned long long (&v)[E], unsigned long long* smem, int tid) {
    constexpr int T = 128, M = E * T;
    // phase A: merges that fit inside one thread
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & j) == 0) ce_regs(v[r], v[r | j], (((tid * E) + r) & k) == 0);
        }
    }
    // phase B: merges spanning threads
    for (int k = 2 * E; k <= M; k <<= 1) {
        const bool asc = ((tid * E) & k) == 0;        // r < E <= k/2 never reaches bit k
        for (int j = k >> 1; j >= E; j >>= 1) {
            const int jl = j / E;
            const bool keep_min = ((tid & jl) == 0) == asc;
            if (jl < 32) {
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], jl);
                    v[r] = keep_min ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < E; r++) smem[r * T + tid] = v[r];      // [r][tid]: conflict-free
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const unsigned long long o = smem[r * T + (tid ^ jl)];
                    v[r] = keep_min ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
                }
                __syncthreads();
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & j) == 0) ce_regs(v[r], v[r | j], asc);
        }
    }
}

template <int E>
__device__ __forceinline__ void sort_tile_regs(unsigned long long* smem, uint32_t tile, uint2 rg,
                                               const unsigned long long* __restrict__ pairs,
                                               uint32_t* __restrict__ point_list,
                                               unsigned long long* __restrict__ keys_sorted, int tid) {
    const int n = (int)(rg.y - rg.x);
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = tid * E + r;
        v[r] = e < n ? pairs[rg.x + e] : kPad;
    }
    block_bitonic_regs<E>(v, smem, tid);
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = tid * E + r;
        if (e < n) {
            point_list[rg.x + e] = (uint32_t)v[r];
            if (keys_sorted) keys_sorted[rg.x + e] = ((unsigned long long)tile << 32) | (v[r] >> 32);
        }
    }
}

// ---- one warp per tile (n <= 512): E keys per lane in registers, every cross-lane stride is a
// shuffle, no shared memory and no block barrier at all; the 8160 warps of a 1080p frame are all
// resident at once.  ~1.7x fewer instructions per tile than the 128-thread version. ----
template <int E>
__device__ __forceinline__ void warp_bitonic_regs(unsigned long long (&v)[E], int lane) {
#pragma unroll
    for (int k = 2; k <= E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & j) == 0) ce_regs(v[r], v[r | j], (((lane * E) + r) & k) == 0);
        }
    }
#pragma unroll
    for (int k = 2 * E; k <= 32 * E; k <<= 1) {
        const bool asc = ((lane * E) & k) == 0;
#pragma unroll
        for (int j = k >> 1; j >= E; j >>= 1) {
            const int jl = j / E;
            const bool keep_min = ((lane & jl) == 0) == asc;
#pragma unroll
            for (int r = 0; r < E; r++) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], jl);
                v[r] = keep_min ? (v[r] < o ? v[r] : o) : (v[r] > o ? v[r] : o);
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; r++)
                if ((r & j) == 0) ce_regs(v[r], v[r | j], asc);
        }
    }
}

template <int E>
__device__ __forceinline__ void sort_tile_warp(uint32_t tile, uint2 rg, const unsigned long long* __restrict__ pairs,
                                               uint32_t* __restrict__ point_list,
                                               unsigned long long* __restrict__ keys_sorted, int lane) {
    const int n = (int)(rg.y - rg.x);
    unsigned long long v[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = lane * E + r;
        v[r] = e < n ? pairs[rg.x + e] : kPad;
    }
    warp_bitonic_regs<E>(v, lane);
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int e = lane * E + r;
        if (e < n) {
            point_list[rg.x + e] = (uint32_t)v[r];
            if (keys_sorted) keys_sorted[rg.x + e] = ((unsigned long long)tile << 32) | (v[r] >> 32);
        }
    }
}

__global__ void __launch_bounds__(128)
tile_sort_warp_kernel(int tiles, const uint2* __restrict__ ranges, const unsigned long long* __restrict__ pairs,
                      uint32_t* __restrict__ point_list, unsigned long long* __restrict__ keys_sorted) {
    const int tile = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= tiles) return;
    const uint2 rg = ranges[tile];
    const int n = (int)(rg.y - rg.x);
    if (n == 0 || n > kWarpMax) return;
    if (n <= 32)       sort_tile_warp<1>(tile, rg, pairs, point_list, keys_sorted, lane);
    else if (n <= 64)  sort_tile_warp<2>(tile, rg, pairs, point_list, keys_sorted, lane);
    else if (n <= 128) sort_tile_warp<4>(tile, rg, pairs, point_list, keys_sorted, lane);
    else if (n <= 256) sort_tile_warp<8>(tile, rg, pairs, point_list, keys_sorted, lane);
    else               sort_tile_warp<16>(tile, rg, pairs, point_list, keys_sorted, lane);
}

// 512 < n <= 2048: 128-thread CTA per tile; persistent CTAs walk the mid-size list.
__global__ void __launch_bounds__(128)
tile_sort_small_kernel(const uint2* __restrict__ ranges, const unsigned long long* __restrict__ pairs,
                       uint32_t* __restrict__ point_list, unsigned long long* __restrict__ keys_sorted,
                       const uint32_t* __restrict__ mid_list, const uint32_t* __restrict__ mid_count) {
    __shared__ unsigned long long s[kSmallMax];
    const uint32_t nm = *mid_count;
    const int tid = threadIdx.x;
    for (uint32_t b = blockIdx.x; b < nm; b += gridDim.x) {
        const uint32_t tile = mid_list[b];
        const uint2 rg = ranges[tile];
        const int n = (int)(rg.y - rg.x);
        if (n <= 1024) sort_tile_regs<8>(s, tile, rg, pairs, point_list, keys_sorted, tid);
        else           sort_tile_regs<16>(s, tile, rg, pairs, point_list, keys_sorted, tid);
        __syncthreads();
    }
}

// Tiles with more than kSmallMax entries: persistent CTAs walk the big-tile list.
__global__ void __launch_bounds__(1024)
tile_sort_large_kernel(const uint2* __restrict__ ranges, unsigned long long* __restrict__ pairs,
                       uint32_t* __restrict__ point_list, unsigned long long* __restrict__ keys_sorted,
                       const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);
    const uint32_t nb = *big_count;
    const int tid = threadIdx.x;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint32_t tile = big_list[b];
        const uint2 rg = ranges[tile];
        const int n = (int)(rg.y - rg.x);
        if (n <= kLargeMax) {
            sort_tile_in_smem<1024>(s, tile, rg, pairs, point_list, keys_sorted, tid);
            __syncthreads();
        } else {
            // global-memory bitonic (same all-ascending network, virtual +inf beyond n)
            unsigned long long* g = pairs + rg.x;
            const long long m = 1ll << (64 - __clzll((long long)n - 1));
            bitonic_ascending<1024, long long>(g, (long long)n, m, tid);
            for (int i = tid; i < n; i += 1024) {
                const unsigned long long v = g[i];
                point_list[rg.x + i] = (uint32_t)v;
                if (keys_sorted) keys_sorted[rg.x + i] = ((unsigned long long)tile << 32) | (v >> 32);
            }
            __syncthreads();
        }
    }
}

size_t bucket_temp_bytes(int tiles) { return align_up((size_t)tiles * 4, 256) * 4 + 512; }

int launch_bucket_binning(int P, size_t R, int gx, int gy, int row0, int row1, const float4* tmat,
                          const int* radii, const uint32_t* offsets, unsigned long long* pairs,
                          uint32_t* point_list, unsigned long long* keys_sorted, uint2* ranges,
                          void* temp, const uint32_t* tile_count_ready, cudaStream_t stream) {
    const int tiles = gx * gy;
    char* c = (char*)temp;
    const size_t stride = align_up((size_t)tiles * 4, 256);
    uint32_t* tile_count = (uint32_t*)c;
    uint32_t* tile_fill = (uint32_t*)(c + stride);
    uint32_t* big_list = (uint32_t*)(c + 2 * stride);
    uint32_t* mid_list = (uint32_t*)(c + 3 * stride);
    uint32_t* big_count = (uint32_t*)(c + 4 * stride);
    uint32_t* mid_count = big_count + 32;
    const int blocks = (P + kWalkBlock - 1) / kWalkBlock;
    if (tile_count_ready) {
        tile_count = const_cast<uint32_t*>(tile_count_ready);   // counted by preprocess_fwd (fused)
    } else {
        SURFEL_CUDA_OK(cudaMemsetAsync(tile_count, 0, (size_t)tiles * 4, stream));
    }
    if (!tile_count_ready && P > 0 && R > 0) {
        LaunchScope scope(kStTileCount, stream);
        tile_count_kernel<<<blocks, kWalkBlock, 0, stream>>>(P, gx, gy, row0, row1, tmat, radii, offsets, tile_count);
        SURFEL_CUDA_OK(cudaGetLastError());
    }
    {
        LaunchScope scope(kStTileScan, stream);
        tile_scan_kernel<<<1, 1024, 0, stream>>>(tiles, (uint32_t)std::min<size_t>(R, 0xffffffffu), tile_count, ranges, tile_fill, big_list, big_count, mid_list, mid_count);
        SURFEL_CUDA_OK(cudaGetLastError());
    }
    if (P <= 0 || R == 0) return 0;
    {
        LaunchScope scope(kStTileScatter, stream);
        tile_scatter_kernel<<<blocks, kWalkBlock, 0, stream>>>(P, gx, gy, row0, row1, tmat, radii, offsets, tile_fill, pairs, (uint32_t)std::min<size_t>(R, 0xffffffffu));
        SURFEL_CUDA_OK(cudaGetLastError());
    }
    {
        LaunchScope scope(kStTileSort, stream);
        tile_sort_warp_kernel<<<(tiles + 3) / 4, 128, 0, stream>>>(tiles, ranges, pairs, point_list, keys_sorted);
        SURFEL_CUDA_OK(cudaGetLastError());
        prof_count_launch();
        tile_sort_small_kernel<<<148 * 4, 128, 0, stream>>>(ranges, pairs, point_list, keys_sorted, mid_list, mid_count);
        SURFEL_CUDA_OK(cudaGetLastError());
    }
    {
        static bool attr_set[kMaxDevices] = {};
        const int slot = current_device_slot();
        if (slot < 0 || !attr_set[slot]) {
            SURFEL_CUDA_OK(cudaFuncSetAttribute(tile_sort_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                kLargeMax * 8));
            if (slot >= 0) attr_set[slot] = true;
        }
        LaunchScope scope(kStTileSort, stream);
        tile_sort_large_kernel<<<148, 1024, kLargeMax * 8, stream>>>(ranges, pairs, point_list, keys_sorted, big_list, big_count);
        SURFEL_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace surfel
