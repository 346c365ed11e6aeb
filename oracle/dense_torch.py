"""Dense pure-PyTorch 2D-surfel rasterizer (every splat x every pixel, global depth sort).

TEST INFRASTRUCTURE ONLY (see oracle/surfel_oracle.c header).  Two jobs:

1. The "pure-Python surfel rasterizer" CPU baseline named by BASELINE.json / SURVEY.md §8(d):
   the authors' Colab notebook (/root/reference/README.md:3) is only a hyperlink, so this is the
   builder-written stand-in with that shape (dense, PyTorch, CPU).
2. An independent cross-check of the C oracle: it is differentiable, so torch.autograd gives
   the gradients the hand-written A.4/A.5 replay in surfel_oracle.c must reproduce (float64).

It follows SURVEY.md Appendix A.1 + A.3 and the in-tree restatements:
  T matrix  /root/reference/gaussian_renderer/__init__.py:64-75, scene/gaussian_model.py:27-33
  rotation  /root/reference/utils/general_utils.py:78-110
  SH        /root/reference/utils/sh_utils.py:57-112 (+0.5 / clamp: gaussian_renderer/__init__.py:91)
Tile faithfulness (a splat only reaches the pixels of the tiles in its rect, A.1 step 6) is kept
by a mask, so results equal the tile-based algorithm exactly, not approximately.

Non-derivative conventions reproduced for the autograd comparison:
  * alpha = min(0.99, opa*G) passes its gradient straight through the clamp;
  * min(rho3d, rho2d) and the depth select differentiate the selected branch only;
  * the rotation gradient is w.r.t. the unit quaternion (no normalisation in the graph);
  * culls, `continue` tests, the T<1e-4 stop and the median pick are non-differentiable masks.
"""
import math

import torch

NEAR_N, FAR_N = 0.2, 100.0
FILTER_INV_SQUARE = 2.0
FILTER_SIZE = 0.707106
CUTOFF = 3.0
BLOCK = 16

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
      0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, d):
    """sh: (P,M,3) coefficient-major; d: (P,3) unit directions -> (P,3)."""
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    r = C0 * sh[:, 0]
    if deg > 0:
        r = r - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6]
             + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        r = (r + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
             + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
             + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
             + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
             + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return r


def quat_to_R(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(-1, 3, 3)


def build_pm(proj, W, H):
    """(4,3): columns x*w, y*w, w of projmatrix @ ndc2pix (pixel-centre convention (W-1)/2)."""
    c0 = proj[:, 0] * (W / 2.0) + proj[:, 3] * ((W - 1) / 2.0)
    c1 = proj[:, 1] * (H / 2.0) + proj[:, 3] * ((H - 1) / 2.0)
    return torch.stack([c0, c1, proj[:, 3]], dim=1)


def preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, W, H,
               sh_degree=3, scale_modifier=1.0, transMat_precomp=None, colors_precomp=None,
               normalize_quat=True):
    """A.1, vectorised.  Returns dict with T (P,9), xy, depth, normal, rgb, radii, rect, visible."""
    P = means3D.shape[0]
    vm, Pm = viewmatrix, build_pm(projmatrix, W, H)
    ones = torch.ones(P, 1, dtype=means3D.dtype)
    p_view = torch.cat([means3D, ones], 1) @ vm[:, :3]
    visible = p_view[:, 2] > NEAR_N
    if transMat_precomp is None:
        q = rotations / rotations.norm(dim=1, keepdim=True) if normalize_quat else rotations
        R = quat_to_R(q)
        L0 = R[:, :, 0] * (scale_modifier * scales[:, 0:1])
        L1 = R[:, :, 1] * (scale_modifier * scales[:, 1:2])
        L2 = R[:, :, 2]
        rows = torch.stack([L0 @ Pm[:3], L1 @ Pm[:3], means3D @ Pm[:3] + Pm[3]], dim=1)  # (P,3i,3j)
        T = rows.permute(0, 2, 1).reshape(P, 9)  # [Tu|Tv|Tw], Tu = (rows[0][0],rows[1][0],rows[2][0])
        normal = L2 @ vm[:3, :3]
    else:
        T = transMat_precomp
        normal = torch.tensor([0.0, 0.0, 1.0], dtype=means3D.dtype).expand(P, 3)
    c = -(p_view[:, :3] * normal).sum(1)
    visible = visible & (c != 0)
    normal = normal * torch.where(c > 0, 1.0, -1.0).to(normal.dtype)[:, None]
    Tu, Tv, Tw = T[:, 0:3], T[:, 3:6], T[:, 6:9]
    t = torch.tensor([CUTOFF * CUTOFF, CUTOFF * CUTOFF, -1.0], dtype=means3D.dtype)
    d = (t * Tw * Tw).sum(1)
    visible = visible & (d != 0)
    dsafe = torch.where(d != 0, d, torch.ones_like(d))
    f = t[None] / dsafe[:, None]
    xy = torch.stack([(f * Tu * Tw).sum(1), (f * Tv * Tw).sum(1)], 1)
    ext2 = xy * xy - torch.stack([(f * Tu * Tu).sum(1), (f * Tv * Tv).sum(1)], 1)
    half = torch.sqrt(torch.clamp_min(ext2, 1e-4))
    radius = torch.ceil(torch.clamp_min(half.max(dim=1).values, CUTOFF * FILTER_SIZE))
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    with torch.no_grad():
        def trunc_div(v):
            return torch.trunc(v / BLOCK).clamp(-2 ** 30, 2 ** 30).long()
        x0 = trunc_div(xy[:, 0] - radius).clamp(0, gx)
        y0 = trunc_div(xy[:, 1] - radius).clamp(0, gy)
        x1 = trunc_div(xy[:, 0] + radius + (BLOCK - 1)).clamp(0, gx)
        y1 = trunc_div(xy[:, 1] + radius + (BLOCK - 1)).clamp(0, gy)
        visible = visible & (((x1 - x0) * (y1 - y0)) > 0)
    if colors_precomp is None:
        dirs = means3D - campos[None]
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        raw = eval_sh(sh_degree, shs, dirs) + 0.5
        clamped = raw < 0
        rgb = torch.clamp_min(raw, 0.0)
    else:
        rgb, clamped = colors_precomp, torch.zeros(P, 3, dtype=torch.bool)
    radii = torch.where(visible, radius.detach(), torch.zeros_like(radius)).to(torch.int64)
    return dict(T=T, xy=xy, depth=p_view[:, 2], normal=normal, rgb=rgb, clamped=clamped,
                radii=radii, rect=torch.stack([x0, y0, x1, y1], 1), visible=visible,
                opacity=opacities.reshape(-1),
                tiles_touched=torch.where(visible, (x1 - x0) * (y1 - y0), torch.zeros_like(x0)))


def rasterize(pre, bg, W, H, pixel_chunk=4096, upstream_lowpass_depth=False, xy_override=None):
    """A.3 dense blend.  Returns color (3,H,W), others (7,H,W), n_contrib (H,W), final_T (H,W).

    upstream_lowpass_depth: emulate, through autograd, the depth gradient the published upstream
      backward applies in the low-pass branch ("Propagate the gradients of depth": dL_dTw += (s.x, s.y, 1)
      * dL_dz although the forward uses depth = Tw.z there): the VALUE stays Tw.z, the gradient is that of
      s.x*Tw.x + s.y*Tw.y + Tw.z with s held fixed.
    xy_override: (P,2) tensor used instead of pre["xy"] for the low-pass distance, e.g. a detached leaf,
      so that T.grad excludes the low-pass filter's gradient (the raw dL_dtransMat upstream's
      densification proxy reads on the scales+rotations path, /root/reference/README.md:118)."""
    dt = pre["T"].dtype
    vis = pre["visible"].nonzero().squeeze(1)
    # global front-to-back order: depth bits ascending, ties by splat index (stable)
    order = vis[torch.argsort(pre["depth"][vis].detach(), stable=True)]
    T9, xy = pre["T"][order], (pre["xy"] if xy_override is None else xy_override)[order]
    Tu, Tv, Tw = T9[:, None, 0:3], T9[:, None, 3:6], T9[:, None, 6:9]
    opa, nrm, rgb = pre["opacity"][order], pre["normal"][order], pre["rgb"][order]
    rect = pre["rect"][order]
    N = W * H
    color = torch.zeros(3, N, dtype=dt)
    others = torch.zeros(7, N, dtype=dt)
    n_contrib = torch.zeros(N, dtype=torch.int64)
    final_T = torch.ones(N, dtype=dt)
    V = order.shape[0]
    if V == 0:
        color = color + bg[:, None]
        return color.reshape(3, H, W), others.reshape(7, H, W), n_contrib.reshape(H, W), final_T.reshape(H, W)
    cs, os_ = [], []
    for s in range(0, N, pixel_chunk):
        e = min(N, s + pixel_chunk)
        pid = torch.arange(s, e)
        px, py = (pid % W), (pid // W)
        pxf, pyf = px.to(dt)[None, :, None], py.to(dt)[None, :, None]
        tx, ty = (px // BLOCK)[None], (py // BLOCK)[None]
        in_tile = ((tx >= rect[:, 0:1]) & (tx < rect[:, 2:3]) & (ty >= rect[:, 1:2]) & (ty < rect[:, 3:4]))
        k = pxf * Tw - Tu  # (V,n,3)
        l = pyf * Tw - Tv
        p = torch.cross(k, l, dim=-1)
        pz = p[..., 2]
        ok = in_tile & (pz != 0)
        pzs = torch.where(pz != 0, pz, torch.ones_like(pz))
        sx, sy = p[..., 0] / pzs, p[..., 1] / pzs
        rho3d = sx * sx + sy * sy
        dx, dy = xy[:, None, 0] - pxf[..., 0], xy[:, None, 1] - pyf[..., 0]
        rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy)
        use3d = rho3d <= rho2d
        rho = torch.where(use3d, rho3d, rho2d)
        depth3d = sx * Tw[..., 0] + sy * Tw[..., 1] + Tw[..., 2]
        depth_lp = Tw[..., 2].expand_as(sx)
        if upstream_lowpass_depth:
            st = sx.detach() * Tw[..., 0] + sy.detach() * Tw[..., 1]
            depth_lp = depth_lp + (st - st.detach())      # value Tw.z, gradient (s.x, s.y, 1) * dL_dz
        depth = torch.where(use3d, depth3d, depth_lp)
        ok = ok & (depth >= NEAR_N)
        power = -0.5 * rho
        ok = ok & ~(power > 0)
        G = torch.exp(torch.where(ok, power, torch.zeros_like(power)))
        a_raw = opa[:, None] * G
        alpha = a_raw + (torch.clamp_max(a_raw, 0.99) - a_raw).detach()  # straight-through clamp
        ok = ok & (alpha >= 1.0 / 255.0)
        a_eff = torch.where(ok, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_incl = torch.cumprod(one_m, dim=0)
        T_before = torch.cat([torch.ones_like(T_incl[:1]), T_incl[:-1]], 0)
        stop = torch.cumsum((ok & (T_incl < 1e-4)).to(torch.int32), 0) > 0
        active = ok & ~stop
        w = torch.where(active, a_eff * T_before, torch.zeros_like(a_eff))
        dsafe = torch.where(active, depth, torch.ones_like(depth))
        m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / dsafe)
        mw, m2w = m * w, m * m * w
        M1b = torch.cumsum(mw, 0) - mw
        M2b = torch.cumsum(m2w, 0) - m2w
        dist = ((m * m * (1 - T_before) + M2b - 2 * m * M1b) * w).sum(0)
        D = (dsafe * w).sum(0)
        Nn = (nrm[:, None, :] * w[..., None]).sum(0)
        C = (rgb[:, None, :] * w[..., None]).sum(0)
        Tf = torch.where(active, one_m, torch.ones_like(one_m)).prod(0)
        med_flag = active & (T_before > 0.5)
        ar = torch.arange(1, V + 1)[:, None]
        med_idx = (med_flag * ar).max(0).values  # 0 = none
        med_depth = torch.where(med_idx > 0,
                                torch.gather(dsafe, 0, (med_idx - 1).clamp_min(0)[None])[0],
                                torch.zeros_like(D))
        contributor = torch.cumsum(in_tile.to(torch.int64), 0)
        n_contrib[s:e] = (contributor * active).max(0).values
        final_T[s:e] = Tf.detach()
        cs.append((C + Tf[:, None] * bg[None]).T)
        os_.append(torch.stack([D, 1 - Tf, Nn[:, 0], Nn[:, 1], Nn[:, 2], med_depth, dist], 0))
    color = torch.cat(cs, 1)
    others = torch.cat(os_, 1)
    return color.reshape(3, H, W), others.reshape(7, H, W), n_contrib.reshape(H, W), final_T.reshape(H, W)


def render(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos, bg, W, H,
           sh_degree=3, scale_modifier=1.0, transMat_precomp=None, colors_precomp=None,
           normalize_quat=True, pixel_chunk=4096, upstream_lowpass_depth=False, detach_lowpass_center=False):
    pre = preprocess(means3D, scales, rotations, opacities, shs, viewmatrix, projmatrix, campos,
                     W, H, sh_degree, scale_modifier, transMat_precomp, colors_precomp,
                     normalize_quat)
    color, others, n_contrib, final_T = rasterize(
        pre, bg, W, H, pixel_chunk, upstream_lowpass_depth,
        pre["xy"].detach() if detach_lowpass_center else None)
    return color, others, pre, n_contrib, final_T


def densification_proxy(pre_T, dL_dT, W, H):
    """A.5 step 5: the (non-derivative) means2D 'gradient' upstream hands to densification."""
    depth = pre_T[:, 8]
    return torch.stack([dL_dT[:, 2] * depth * 0.5 * W, dL_dT[:, 5] * depth * 0.5 * H], 1)


def fov_pair(fovy_deg, W, H):
    tanfovy = math.tan(math.radians(fovy_deg) / 2)
    return tanfovy * W / H, tanfovy
