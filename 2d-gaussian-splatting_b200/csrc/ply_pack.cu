// ply_pack.cu — SURVEY §8(f) row f4: the on-disk model format either side of the path.
//
// The reference stores a trained model as a binary little-endian PLY with one 61-float row per splat
// (/root/reference/scene/gaussian_model.py:176-209): x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity
// scale_0..1 rot_0..3, all PRE-activation, SH coefficients channel-major (f_rest_{c*15+k}).  The
// rasterizer wants activated, AoS inputs: means3D (P,3), shs (P,16,3) coefficient-major, opacity =
// sigmoid, scales = exp, rotations = normalised quaternion (gaussian_model.py:35-41, :95-115).
//
//  * surfel_ply_unpack: raw rows (already in device memory) -> op inputs (activate = 1) or the
//    reference's parameter tensors (activate = 0), in ONE pass: each warp stages 32 rows through
//    shared memory so that both the 244-byte-stride reads and the five output streams are coalesced;
//    the column of every target float comes from a 61-entry table, so any property order a PLY header
//    declares is handled (load_ply addresses properties by NAME).
//  * surfel_ply_pack: the inverse for save_ply (parameters -> rows in the reference's column order,
//    normals zero).
// HBM-streaming byte shuffling: 244 B in + 232 B out per splat.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/surfel_rasterizer.h"
#include "common.cuh"
#include "profile.h"

namespace surfel {

constexpr int kPlyTargets = SURFEL_PLY_TARGETS;    // 58: xyz 3, sh 48, opacity 1, scale 2, rot 4 (normals dropped)
constexpr int kPlyMaxRow = SURFEL_PLY_MAX_ROW_FLOATS;

struct PlyTable { int col[kPlyTargets]; };

// target order: 0..2 xyz | 3..50 shs as [k][c] (k = coefficient 0..15, c = channel) | 51 opacity |
// 52..53 scale | 54..57 rot
__global__ void __launch_bounds__(128) ply_unpack_kernel(int P, int row_floats, const float* __restrict__ rows,
                                                         const __grid_constant__ PlyTable t, int activate,
                                                         float* __restrict__ means3D, float* __restrict__ shs,
                                                         float* __restrict__ opacities, float* __restrict__ scales,
                                                         float* __restrict__ rotations) {
    extern __shared__ float s_rows[];               // 4 warps x 32 rows x (row_floats | 1)
    __shared__ int s_col[kPlyTargets];
    if (threadIdx.x < kPlyTargets) s_col[threadIdx.x] = t.col[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = row_floats | 1;              // odd: conflict-free column access
    float* panel = s_rows + (size_t)warp * 32 * stride;
    const int first = (blockIdx.x * 4 + warp) * 32;
    if (first >= P) return;
    const int nrows = min(32, P - first);
    const float* src = rows + (size_t)first * row_floats;
    for (int r = 0; r < nrows; r++)                             // rows are contiguous: coalesced, no division
        for (int c = lane; c < row_floats; c += 32) panel[r * stride + c] = __ldg(src + (size_t)r * row_floats + c);
    __syncwarp();
    // stream out, one output tensor at a time, consecutive lanes -> consecutive floats
    for (int f = lane; f < nrows * 3; f += 32) {
        const int r = f / 3, c = f - r * 3;
        means3D[(size_t)first * 3 + f] = panel[r * stride + s_col[c]];
    }
    for (int f = lane; f < nrows * 48; f += 32) {
        const int r = f / 48, c = f - r * 48;
        shs[(size_t)first * 48 + f] = panel[r * stride + s_col[3 + c]];
    }
    if (lane < nrows) {
        const float* row = panel + lane * stride;
        float o = row[s_col[51]], s0 = row[s_col[52]], s1 = row[s_col[53]];
        float qw = row[s_col[54]], qx = row[s_col[55]], qy = row[s_col[56]], qz = row[s_col[57]];
        if (activate) {
            o = 1.0f / (1.0f + expf(-o));                                    // torch.sigmoid
            s0 = expf(s0); s1 = expf(s1);                                    // torch.exp
            const float n = fmaxf(sqrtf(qw * qw + qx * qx + qy * qy + qz * qz), 1e-12f);   // F.normalize eps
            qw /= n; qx /= n; qy /= n; qz /= n;
        }
        const size_t i = (size_t)first + lane;
        opacities[i] = o;
        scales[2 * i] = s0; scales[2 * i + 1] = s1;
        reinterpret_cast<float4*>(rotations)[i] = make_float4(qw, qx, qy, qz);
    }
}

__global__ void __launch_bounds__(128) ply_pack_kernel(int P, const float* __restrict__ xyz, const float* __restrict__ f_dc,
                                                       const float* __restrict__ f_rest, const float* __restrict__ opacity,
                                                       const float* __restrict__ scaling, const float* __restrict__ rotation,
                                                       float* __restrict__ rows) {
    // reference column order: x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..1 rot_0..3 (61 floats)
    __shared__ float s_rows[4 * 32 * 63];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* panel = s_rows + warp * 32 * 63;
    const int first = (blockIdx.x * 4 + warp) * 32;
    if (first >= P) return;
    const int nrows = min(32, P - first);
    for (int f = lane; f < nrows * 3; f += 32) {
        const int r = f / 3, c = f - r * 3;
        panel[r * 63 + c] = xyz[(size_t)first * 3 + f];
        panel[r * 63 + 3 + c] = 0.0f;                                        // normals: zeros_like(xyz)
        panel[r * 63 + 6 + c] = f_dc[(size_t)first * 3 + f];                 // (P,1,3) -> f_dc_c
    }
    for (int f = lane; f < nrows * 45; f += 32) {                            // (P,15,3)[k][c] -> f_rest_{c*15+k}
        const int r = f / 45, kc = f - r * 45, k = kc / 3, c = kc - 3 * k;
        panel[r * 63 + 9 + c * 15 + k] = f_rest[(size_t)first * 45 + f];
    }
    if (lane < nrows) {
        const size_t i = (size_t)first + lane;
        float* row = panel + lane * 63;
        row[54] = opacity[i];
        row[55] = scaling[2 * i]; row[56] = scaling[2 * i + 1];
        const float4 q = reinterpret_cast<const float4*>(rotation)[i];
        row[57] = q.x; row[58] = q.y; row[59] = q.z; row[60] = q.w;
    }
    __syncwarp();
    float* dst = rows + (size_t)first * SURFEL_PLY_ROW_FLOATS;
    for (int f = lane; f < nrows * SURFEL_PLY_ROW_FLOATS; f += 32) {
        const int r = f / SURFEL_PLY_ROW_FLOATS, c = f - r * SURFEL_PLY_ROW_FLOATS;
        dst[f] = panel[r * 63 + c];
    }
}

}  // namespace surfel

using namespace surfel;

extern "C" {

int surfel_ply_unpack(int P, int row_floats, const float* rows, const int32_t* columns, int activate,
                      float* means3D, float* shs, float* opacities, float* scales, float* rotations, void* stream) {
    if (P < 0) { surfel_set_error("surfel_ply_unpack: P < 0"); return 1; }
    if (P == 0) return 0;
    if (row_floats < 1 || row_floats > kPlyMaxRow) {
        surfel_set_error("surfel_ply_unpack: row of %d floats unsupported (max %d)", row_floats, kPlyMaxRow);
        return 1;
    }
    if (!rows || !columns || !means3D || !shs || !opacities || !scales || !rotations) {
        surfel_set_error("surfel_ply_unpack: NULL pointer");
        return 1;
    }
    if (reinterpret_cast<uintptr_t>(rotations) % 16 != 0) { surfel_set_error("surfel_ply_unpack: rotations must be 16-byte aligned"); return 1; }
    PlyTable t;
    for (int i = 0; i < kPlyTargets; i++) {
        if (columns[i] < 0 || columns[i] >= row_floats) {
            surfel_set_error("surfel_ply_unpack: column %d of target %d outside the %d-float row", columns[i], i, row_floats);
            return 1;
        }
        t.col[i] = columns[i];
    }
    const size_t smem = (size_t)4 * 32 * (row_floats | 1) * sizeof(float);
    if (smem > 48 * 1024)       // rows wider than 95 floats (never the reference's 61): opt in per launch; the attribute is per device
        SURFEL_CUDA_OK(cudaFuncSetAttribute(ply_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LaunchScope scope(kStPlyUnpack, (cudaStream_t)stream);
    ply_unpack_kernel<<<(P + 127) / 128, 128, smem, (cudaStream_t)stream>>>(P, row_floats, rows, t, activate, means3D,
                                                                           shs, opacities, scales, rotations);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

int surfel_ply_pack(int P, const float* xyz, const float* features_dc, const float* features_rest,
                    const float* opacity, const float* scaling, const float* rotation, float* rows, void* stream) {
    if (P < 0) { surfel_set_error("surfel_ply_pack: P < 0"); return 1; }
    if (P == 0) return 0;
    if (!xyz || !features_dc || !features_rest || !opacity || !scaling || !rotation || !rows) {
        surfel_set_error("surfel_ply_pack: NULL pointer");
        return 1;
    }
    if (reinterpret_cast<uintptr_t>(rotation) % 16 != 0) { surfel_set_error("surfel_ply_pack: rotation must be 16-byte aligned"); return 1; }
    LaunchScope scope(kStPlyPack, (cudaStream_t)stream);
    ply_pack_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, xyz, features_dc, features_rest, opacity,
                                                                      scaling, rotation, rows);
    SURFEL_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
