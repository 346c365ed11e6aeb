// render_common.cuh — per-(pixel,splat) ray–splat evaluation shared by render forward/backward.
//
// Follows SURVEY.md Appendix A.3 (upstream renderCUDA; source not in /root/reference).  The
// arithmetic up to `alpha` is written with explicit round-to-nearest intrinsics so that forward and
// backward (separate translation units) take bit-identical skip/contribute decisions for a pair —
// the backward replay of transmittance depends on that.
#pragma once
#include "common.cuh"

namespace surfel {

struct PairEval {
    float kx, ky, kz, lx, ly, lz;   // k = px*Tw - Tu, l = py*Tw - Tv
    float pz, inv_pz;               // cross(k,l).z and its reciprocal
    float sx, sy;                   // splat-space intersection
    float dx, dy;                   // xy - pixel
    float depth, G, alpha;
    bool use3d;                     // rho3d <= rho2d (ray-splat branch), else low-pass branch
};

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// q0 = (Tu.x,Tu.y,Tu.z,Tv.x) q1 = (Tv.y,Tv.z,Tw.x,Tw.y) q2 = (Tw.z, xy.x, xy.y, opacity)
// Returns false when one of the A.3 `continue` tests that precede the transmittance test fires.
__device__ __forceinline__ bool eval_pair(float pxf, float pyf, const float4& q0, const float4& q1,
                                          const float4& q2, PairEval& e) {
    const float Twx = q1.z, Twy = q1.w, Twz = q2.x;
    e.kx = __fmaf_rn(pxf, Twx, -q0.x); e.ky = __fmaf_rn(pxf, Twy, -q0.y); e.kz = __fmaf_rn(pxf, Twz, -q0.z);
    e.lx = __fmaf_rn(pyf, Twx, -q0.w); e.ly = __fmaf_rn(pyf, Twy, -q1.x); e.lz = __fmaf_rn(pyf, Twz, -q1.y);
    const float ppx = __fmaf_rn(e.ky, e.lz, -__fmul_rn(e.kz, e.ly));
    const float ppy = __fmaf_rn(e.kz, e.lx, -__fmul_rn(e.kx, e.lz));
    e.pz = __fmaf_rn(e.kx, e.ly, -__fmul_rn(e.ky, e.lx));
    // The A.3 `continue` tests are folded into one predicate instead of four early exits: a warp
    // practically never fails one of the first three on all 32 lanes, so as branches they only cost
    // issue slots and branch-resolve stalls.  Lanes that fail keep computing on inf/NaN, harmlessly.
    bool ok = e.pz != 0.0f;
    const float inv = fast_rcp(e.pz);
    e.inv_pz = inv;
    e.sx = __fmul_rn(ppx, inv); e.sy = __fmul_rn(ppy, inv);
    const float rho3d = __fmaf_rn(e.sx, e.sx, __fmul_rn(e.sy, e.sy));
    e.dx = __fsub_rn(q2.y, pxf); e.dy = __fsub_rn(q2.z, pyf);
    const float rho2d = __fmul_rn(kFilterInvSquare, __fmaf_rn(e.dx, e.dx, __fmul_rn(e.dy, e.dy)));
    e.use3d = rho3d <= rho2d;
    const float rho = fminf(rho3d, rho2d);
    e.depth = e.use3d ? __fadd_rn(__fmaf_rn(e.sx, Twx, __fmul_rn(e.sy, Twy)), Twz) : Twz;
    ok &= !(e.depth < kNear);
    const float power = __fmul_rn(-0.5f, rho);
    ok &= !(power > 0.0f);
    e.G = fast_ex2(__fmul_rn(power, 1.4426950408889634f));
    e.alpha = fminf(kAlphaMax, __fmul_rn(q2.w, e.G));
    ok &= !(e.alpha < kAlphaMin);
    return ok;
}

// Position of the highest set bit (FLO) and the mask of the bits below a position (BMSK): the hit
// loops of the render kernels peel ballot bits with exactly these two instructions.
__device__ __forceinline__ uint32_t high_bit(uint32_t m) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
}
__device__ __forceinline__ uint32_t low_mask(uint32_t width) {
    uint32_t r;
    asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r) : "r"(width));
    return r;
}

// Explicit shared-window addressing: one cvta per kernel instead of a generic->shared conversion
// (S2UR CgaCtaId + ULEA) re-materialised in every inner-loop iteration.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float a) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Warp footprint inside a 16x16 tile: 8 (x) by 4 (y) pixels; warp w sits at (w&1, w>>1).
__device__ __forceinline__ void warp_pixel(int warp, int lane, int& lx, int& ly) {
    lx = ((warp & 1) << 3) + (lane & 7);
    ly = ((warp >> 1) << 2) + (lane >> 3);
}

}  // namespace surfel
