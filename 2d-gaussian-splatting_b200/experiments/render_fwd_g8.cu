// render_fwd_g8.cu — forward blend, "four 8-lane groups per warp" variant.
//
// Same algorithm and staging as render_fwd.cu, different work mapping.  With the
// 8x4 warp footprint only ~9 of 32 lanes contribute to a hit splat (profiles/r1): instruction issue,
// the bound of this kernel, is spent on idle lanes.  Here a warp is four independent 8-lane groups,
// each owning a 4x2 pixel block and walking ITS OWN compacted hit list in lock-step with the other
// groups (same instruction stream, different splat per group).  A 4x2 footprint is hit by ~45 splats
// per tile against ~73 for 8x4, and the longest of the four lists averages ~49, so the number of
// evaluate+blend rounds per warp drops by ~1.5x (simulated on the headline scene).
//
// Per batch of 256 staged splats every group tests all of them against its footprint, eight at a
// time (one lane per splat, bbox in record quad 5), and compacts the hits with ballot + popc into a
// byte queue in shared memory; the blend loop then runs max(queue lengths) rounds.  The bbox test is
// the same conservative one as before, so skipping is exact.
#include "render_common.cuh"
#include "kernels.h"
#include "profile.h"

namespace surfel {

constexpr int kBatchG = 256;

__global__ void __launch_bounds__(256) render_fwd_g8_kernel(RenderParams p) {
    __shared__ float4 s_rec[kRecQuads * kBatchG];        // [quad][slot]
    __shared__ uint8_t s_queue[8 * 4 * kBatchG];         // [warp][group][position] -> slot

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 3, l8 = lane & 7;
    const int tx = blockIdx.x, ty = blockIdx.y + p.row0;
    // group footprint: 4 (x) by 2 (y) pixels inside the warp's 8x4 block
    const int gx0 = tx * kBlockX + ((warp & 1) << 3) + ((grp & 1) << 2);
    const int gy0 = ty * kBlockY + ((warp >> 1) << 2) + ((grp >> 1) << 1);
    const int px = gx0 + (l8 & 3), py = gy0 + (l8 >> 2);
    const bool inside = px < p.W && py < p.H;
    const float pxf = (float)px, pyf = (float)py;
    const float fx0 = (float)gx0, fx1 = fx0 + 3.0f, fy0 = (float)gy0, fy1 = fy0 + 1.0f;

    const uint2 range = p.ranges[ty * p.gx + tx];
    const int total = (int)(range.y - range.x);
    const uint32_t rec_base = smem_u32(s_rec);
    uint8_t* queue = s_queue + (warp * 4 + grp) * kBatchG;
    const unsigned l8_lt = (1u << l8) - 1u;
    constexpr float kMScale = kFar / (kFar - kNear);

    float T = 1.0f, C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0, D = 0, M1 = 0, M2 = 0, dist = 0;
    float median_depth = 0;
    uint32_t last_contributor = 0, median_contributor = 0xFFFFFFFFu;
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);

    for (int base = 0; base < total; base += kBatchG) {
        if (!__syncthreads_or(!warp_done)) break;
        const int n = min(kBatchG, total - base);
        if (tid < n) {
            const uint32_t id = p.point_list[range.x + base + tid];
            const float4* r = p.rec + (size_t)id * kRecQuads;
#pragma unroll
            for (int q = 0; q < kRecQuads; q++) s_rec[q * kBatchG + tid] = __ldg(r + q);
        }
        __syncthreads();
        if (warp_done) continue;

        // ---- per-group hit queues (order preserved: front to back) ----
        int cnt = 0;
        for (int c = 0; c < n; c += 8) {
            const int slot = c + l8;
            bool hit = false;
            if (slot < n) {
                const float4 bb = lds128(rec_base + (5 * kBatchG + slot) * 16);
                hit = bb.x <= fx1 && bb.z >= fx0 && bb.y <= fy1 && bb.w >= fy0;
            }
            const unsigned seg = (__ballot_sync(0xffffffffu, hit) >> (grp << 3)) & 0xFFu;
            if (hit) queue[cnt + __popc(seg & l8_lt)] = (uint8_t)slot;
            cnt += __popc(seg);
        }
        int rounds = cnt;
        rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, 8));
        rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, 16));
        __syncwarp();

        // ---- evaluate + blend, one queue entry per group and round ----
        for (int i = 0; i < rounds; i++) {
            if (i < cnt && !done) {
                const int k = queue[i];
                const uint32_t ra = rec_base + k * 16;
                const float4 q0 = lds128(ra), q1 = lds128(ra + kBatchG * 16), q2 = lds128(ra + 2 * kBatchG * 16);
                PairEval e;
                if (eval_pair(pxf, pyf, q0, q1, q2, e)) {
                    const float test_T = T * (1.0f - e.alpha);
                    if (test_T < kTMin) {
                        done = true;
                    } else {
                        const uint32_t contributor = (uint32_t)(base + k + 1);
                        const float4 q3 = lds128(ra + 3 * kBatchG * 16), q4 = lds128(ra + 4 * kBatchG * 16);
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
            }
            if ((i & 7) == 7 && __all_sync(0xffffffffu, done || i + 1 >= cnt)) break;
        }
        warp_done = __all_sync(0xffffffffu, done);
        __syncwarp();
    }

    if (inside) {
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)py * p.W + px;
        p.accum[pix] = T; p.accum[HW + pix] = M1; p.accum[2 * HW + pix] = M2;
        p.n_contrib[pix] = last_contributor; p.n_contrib[HW + pix] = median_contributor;
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

int launch_render_fwd_g8(const RenderParams& p, cudaStream_t stream) {
    const int rows = p.row1 - p.row0;
    if (rows <= 0 || p.gx <= 0) return 0;
    dim3 grid(p.gx, rows);
    LaunchScope scope(kStRenderFwd, stream);
    render_fwd_g8_kernel<<<grid, 256, 0, stream>>>(p);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace surfel
