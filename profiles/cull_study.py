"""Round-2 study (CPU only): how much would an exact footprint test save over the AABB test?

The render kernels cull a splat for a warp's 8x4 pixel footprint with the splat's conservative screen
AABB (record quad 5).  ncu says 20 % of the (warp, splat) pairs that pass are dead: no pixel of the
footprint reaches alpha >= 1/255.  Candidate replacement, evaluated here in float32 exactly as a kernel
would: the footprint rectangle against  {rho3d <= tau} (an ellipse  d^T M d <= 1 around e)  UNION  the
low-pass disk {2 |p - c|^2 <= tau}:
    min over the rectangle of the quadratic form = min over the two edges facing e (closed form),
    disk: squared distance from c to the rectangle.
Reports, on a sample of visible splats of a BASELINE config: pairs passing the AABB test, pairs that are
live, pairs passing the exact test, and — the number that matters — live pairs the exact test would
WRONGLY cull (must be 0 with the margins used).

  python profiles/cull_study.py [workload] [sample]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "2d-gaussian-splatting_b200"))
import surfel_scenes as S
from oracle import surfel_oracle as O

f32 = np.float32
MARGIN_Q = f32(1.05)        # ellipse threshold (exact: 1)
MARGIN_R = f32(0.05)        # added to the disk radius, pixels


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "headline"
    nsample = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    scene, cam = S.named(workload)
    P, W, H = S.CONFIGS[workload]
    sc, cm = S.to_numpy(scene), S.to_numpy(cam)
    pre = O.preprocess_fwd(sc["means3D"], sc["scales"], sc["rotations"], sc["opacities"], sc["shs"], cm["viewmatrix"],
                           cm["projmatrix"], cm["campos"], W, H)
    vis = np.nonzero(pre["radii"] > 0)[0]
    rng = np.random.default_rng(0)
    samp = rng.choice(vis, min(nsample, len(vis)), replace=False)
    n_aabb = n_live = n_exact = n_wrong = n_unbounded = 0
    for i in samp:
        tm = pre["transMat"][i].astype(f32)
        cx, cy = (f32(v) for v in pre["xy"][i])
        opa = f32(pre["normal_opacity"][i, 3])
        a255 = f32(255.0) * opa
        if a255 < f32(0.999):
            continue
        tau = f32(2.0) * np.log(a255, dtype=f32) + f32(0.01)
        r = np.sqrt(f32(0.5) * tau, dtype=f32) + MARGIN_R
        Tu, Tv, Tw = tm[0:3], tm[3:6], tm[6:9]
        # pp(x, y) = a x + b y + c ; rho3d <= tau  <=>  ppx^2 + ppy^2 - tau ppz^2 <= 0
        a, b, c = np.cross(Tv, Tw).astype(f32), np.cross(Tw, Tu).astype(f32), np.cross(Tu, Tv).astype(f32)
        c = a * cx + b * cy + c            # conic about the splat's own screen position: without this shift the
        #                                    centre and Q(centre) cancel 1e6-sized float32 terms (pixel coordinates squared)
        wgt = np.array([1, 1, -tau], dtype=f32)
        M00, M01, M11 = np.sum(wgt * a * a), np.sum(wgt * a * b), np.sum(wgt * b * b)
        m0, m1, m22 = np.sum(wgt * a * c), np.sum(wgt * b * c), np.sum(wgt * c * c)
        det = M00 * M11 - M01 * M01
        # same population the kernel culls today (preprocess_fwd.cu: bounded conic, splat in front)
        wxy, wz2 = Tw[0] * Tw[0] + Tw[1] * Tw[1], Tw[2] * Tw[2]
        bounded = Tw[2] > 0 and wz2 > f32(1.05) * tau * wxy and det > 0 and M00 > 0
        if bounded:
            ex = (M01 * m1 - M11 * m0) / det                  # relative to (cx, cy)
            ey = (M01 * m0 - M00 * m1) / det
            qe = m22 + m0 * ex + m1 * ey                      # Q at the centre (< 0 inside)
            bounded = qe < 0
            ex, ey = cx + ex, cy + ey
        if bounded:
            A, B, C = M00 / -qe, M01 / -qe, M11 / -qe
            hx, hy = np.sqrt(C / (A * C - B * B)), np.sqrt(A / (A * C - B * B))      # AABB half extents
            bx0, bx1 = min(cx - r, ex - hx - f32(0.05)), max(cx + r, ex + hx + f32(0.05))
            by0, by1 = min(cy - r, ey - hy - f32(0.05)), max(cy + r, ey + hy + f32(0.05))
            bounded = bool(np.isfinite([bx0, bx1, by0, by1]).all())
        if not bounded:
            n_unbounded += 1
            continue                                          # never culled today either
        x0, x1 = max(0, int(np.ceil(bx0))), min(W - 1, int(np.floor(bx1)))
        y0, y1 = max(0, int(np.ceil(by0))), min(H - 1, int(np.floor(by1)))
        if x1 < x0 or y1 < y0:
            continue
        X0, Y0 = (x0 // 8) * 8, (y0 // 4) * 4
        xs = np.arange(X0, min((x1 // 8) * 8 + 7, W - 1) + 1, dtype=f32)
        ys = np.arange(Y0, min((y1 // 4) * 4 + 3, H - 1) + 1, dtype=f32)
        px, py = np.meshgrid(xs, ys)
        kx, ky, kz = px * Tw[0] - Tu[0], px * Tw[1] - Tu[1], px * Tw[2] - Tu[2]
        lx, ly, lz = py * Tw[0] - Tv[0], py * Tw[1] - Tv[1], py * Tw[2] - Tv[2]
        ppx, ppy, pz = ky * lz - kz * ly, kz * lx - kx * lz, kx * ly - ky * lx
        with np.errstate(all="ignore"):
            sx, sy = ppx / pz, ppy / pz
            rho = np.minimum(sx * sx + sy * sy, f32(2.0) * ((cx - px) ** 2 + (cy - py) ** 2))
            alpha = np.minimum(f32(0.99), opa * np.exp(f32(-0.5) * rho))
        live = alpha >= f32(1.0 / 255.0) * f32(0.999)          # slack for rcp.approx / ex2.approx
        for r0 in range(0, len(ys), 4):
            for c0 in range(0, len(xs), 8):
                fx0, fx1 = xs[c0], xs[min(c0 + 7, len(xs) - 1)]
                fy0, fy1 = ys[r0], ys[min(r0 + 3, len(ys) - 1)]
                if not (bx0 <= fx1 and bx1 >= fx0 and by0 <= fy1 and by1 >= fy0):
                    continue
                n_aabb += 1
                is_live = bool(live[r0:r0 + 4, c0:c0 + 8].any())
                n_live += is_live
                # --- the exact test, float32 ---
                dx0, dx1, dy0, dy1 = fx0 - ex, fx1 - ex, fy0 - ey, fy1 - ey
                xc, yc = min(max(f32(0), dx0), dx1), min(max(f32(0), dy0), dy1)
                dys = min(max(-B * xc / C, dy0), dy1)
                dxs = min(max(-B * yc / A, dx0), dx1)
                q1 = A * xc * xc + f32(2) * B * xc * dys + C * dys * dys
                q2 = A * dxs * dxs + f32(2) * B * dxs * yc + C * yc * yc
                ddx, ddy = min(max(cx, fx0), fx1) - cx, min(max(cy, fy0), fy1) - cy
                passes = min(q1, q2) <= MARGIN_Q or ddx * ddx + ddy * ddy <= r * r
                n_exact += passes
                n_wrong += (is_live and not passes)
                if is_live and not passes and os.environ.get("CULL_DEBUG"):
                    print("WRONG", i, float(min(q1, q2)), float(ddx * ddx + ddy * ddy), float(r * r), float(A), float(B), float(C),
                          float(ex), float(ey), float(cx), float(cy), float(alpha[r0:r0 + 4, c0:c0 + 8].max()), file=sys.stderr)
    out = {"workload": workload, "sampled_visible_splats": int(len(samp)), "unbounded_conics_skipped": int(n_unbounded),
           "pairs_passing_aabb": int(n_aabb), "pairs_live": int(n_live), "pairs_passing_exact_test": int(n_exact),
           "live_pairs_wrongly_culled": int(n_wrong),
           "dead_fraction_today": 1 - n_live / max(1, n_aabb), "dead_fraction_with_exact_test": 1 - n_live / max(1, n_exact),
           "evaluations_saved": 1 - n_exact / max(1, n_aabb)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
