"""Parity bars shared by the GPU tests, and the error log (gpurun_out/parity_stats.jsonl)."""
import json
import os

import numpy as np

# Small scenes (<= 333 px): the float32 oracle is still within ~1e-6 of the exact value, so the plain
# comparison against it carries the bar: 1e-4 per pixel, with the share of pixels that may flip one of the
# discontinuous tests (alpha < 1/255, T < 1e-4, rho3d <= rho2d, T > 0.5) cut to what the data support
# (worst observed 5.3e-5; see DESIGN.md section 2 for the table).
FLIP_BUDGET = 4e-4


def record_stats(name, err, extra=None):
    """Print and log (gpurun_out/parity_stats.jsonl) the error distribution of one tensor: the worst
    offender, the 99.9th and 99th percentiles and the median — what the tolerances below are cut to."""
    import json, os
    err = np.asarray(err, np.float64).ravel()
    fin = err[np.isfinite(err)]
    st = dict(name=name, n=int(err.size), nonfinite=int(err.size - fin.size),
              max=float(fin.max()) if fin.size else 0.0,
              p999=float(np.quantile(fin, 0.999)) if fin.size else 0.0,
              p99=float(np.quantile(fin, 0.99)) if fin.size else 0.0,
              p50=float(np.quantile(fin, 0.5)) if fin.size else 0.0)
    st.update(extra or {})
    st["test"] = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    print(f"{name}: max {st['max']:.3e}  p99.9 {st['p999']:.3e}  p99 {st['p99']:.3e}  median {st['p50']:.3e}  (n={st['n']})")
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps(st) + "\n")
    except OSError:
        pass
    return st


def assert_close_budget(name, got, ref, tol=1e-4, budget=FLIP_BUDGET):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    bad = (err > tol) | ~np.isfinite(got)
    frac = bad.mean()
    record_stats(name, err, dict(tol=tol, outside=int(bad.sum()), frac_outside=float(frac), budget=budget))
    assert frac <= budget, f"{name}: {frac:.3e} of entries outside {tol} (budget {budget})"
    return frac



def grad_check(name, got, ref, rtol=2e-3, budget=3e-3):
    """Per-splat gradient rows: error relative to the row's own magnitude plus a floor tied to the
    tensor's scale (sums of O(100) float32 atomics in a different order than the oracle's double sum)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    got = got.reshape(ref.shape)
    scale = np.abs(ref).max() + 1e-30
    err = np.abs(got - ref) / (np.abs(ref) + 1e-3 * scale)
    bad = (err > rtol) | ~np.isfinite(got)
    record_stats(name, err, dict(tol=rtol, outside=int(bad.sum()), frac_outside=float(bad.mean()), budget=budget, ref_max=float(scale)))
    assert bad.mean() <= budget, f"{name}: {bad.mean():.3e} outside tolerance"



# Bars.  The reference computes in float32, and its intersection k = px*Tw - Tu cancels terms of order
# |pixel| * depth: at 1080p-8K its own float32 result (restated by the float32 oracle, or upstream's
# FMA-contracted build) is only accurate to ~1e-5..1e-4, which flips the alpha >= 1/255 / T < 1e-4 / rho3d <=
# rho2d / T > 0.5 decisions on a measurable fraction of pixels.  The CUDA path evaluates the same formulas
# about the splat's own screen position (common.cuh) and lands within a few 1e-7 of their EXACT value.
# So every tensor is held to two bars:
#   (1) against the exact (float64) evaluation of the published formulas on the same float32 inputs:
#       the fraction of entries outside the tolerance must stay below a tight budget;
#   (2) against the float32 oracle: the fraction outside the tolerance may not exceed what the float32
#       oracle's own distance to the exact value explains (NOISE_FACTOR x that fraction + a small floor).
# budgets cut to the data of the round-2 runs (worst observed share outside: outputs 1.9e-6, gradients 2.3e-5;
# p99.9 of the error: outputs 4e-7, gradients 2.3e-5 — DESIGN.md section 2 has the table)
OUT_TOL, OUT_BUDGET_EXACT = 1e-4, 1e-5           # per-pixel outputs vs exact: 1e-4 on all but 1e-5 of the entries
GRAD_TOL, GRAD_BUDGET_EXACT = 5e-4, 1e-4         # gradients vs exact (error relative to |ref| + 1e-3 max|ref|)
NOISE_FACTOR, NOISE_FLOOR = 1.5, 2e-4


def rel_out(a, b):
    return np.abs(a.astype(np.float64) - b) / np.maximum(1.0, np.abs(b))


def rel_grad(a, b):
    a, b = np.asarray(a, np.float64).reshape(np.asarray(b).shape), np.asarray(b, np.float64)
    return np.abs(a - b) / (np.abs(b) + 1e-3 * (np.abs(b).max() + 1e-30))


def two_bar_check(name, got, ref32, ref64, rel, tol, budget_exact, p999_bar=None):
    """p999_bar: bound on the 99.9th percentile of the distance to the exact value (default tol / 4; gradients
    under the distortion-dominated regularizer cotangents of config 3 use 0.7 tol — their per-splat sums cancel, and
    the float32 atomics' rounding is relative to the terms, not to the sum; worst observed 2.5e-4 = 0.5 tol)."""
    p999_bar = tol / 4 if p999_bar is None else p999_bar
    e_exact, e_f32, e_noise = rel(got, ref64), rel(got, ref32), rel(ref32, ref64)
    assert np.isfinite(np.asarray(got)).all(), name
    s_exact = record_stats(f"{name}: CUDA vs float64 evaluation", e_exact, dict(tol=tol, frac_outside=float((e_exact > tol).mean()), budget=budget_exact))
    s_f32 = record_stats(f"{name}: CUDA vs float32 oracle", e_f32, dict(tol=tol, frac_outside=float((e_f32 > tol).mean())))
    s_noise = record_stats(f"{name}: float32 oracle vs float64 evaluation", e_noise, dict(tol=tol, frac_outside=float((e_noise > tol).mean())))
    assert s_exact["frac_outside"] <= budget_exact, f"{name}: {s_exact['frac_outside']:.2e} of entries further than {tol} from the exact value"
    assert s_exact["p999"] <= p999_bar, f"{name}: p99.9 of the distance to the exact value is {s_exact['p999']:.2e} (bar {p999_bar:.2e})"
    allowed = NOISE_FACTOR * s_noise["frac_outside"] + NOISE_FLOOR
    assert s_f32["frac_outside"] <= allowed, \
        f"{name}: {s_f32['frac_outside']:.2e} outside {tol} vs the float32 oracle; its own rounding noise explains {allowed:.2e}"


