"""Tolerance rules shared by the parity tests.

north_star: indexing / integer work bit-exact; fp32 losses and gradients within 1e-5 relative.
"Relative" is taken norm-wise for tensors (|a-b| <= rtol * max|ref|): element-wise relative error is
undefined where a gradient crosses zero.  Losses that contain |.| (sad, census_sad, smoothness) have a
sign() in their gradient: where the argument is within rounding of zero two correct fp32 evaluations
may pick different signs, so a vanishing fraction of elements (<= outlier_frac) may exceed rtol.
"""
import json
import os

import numpy as np

RTOL = 1e-5


def _record(kind, name, measured, allowed):
    """Every parity assertion logs what it MEASURED next to what it allows (DIS_PARITY_REPORT=<file>.jsonl; printed with
    pytest -s).  profiles/r02_parity_report.jsonl is one such run: a tolerance is a claim only beside its measurement."""
    rec = {"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "check": name, "kind": kind, **measured, **allowed}
    print("PARITY", json.dumps(rec))
    path = os.environ.get("DIS_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")


def to_np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def max_rel(a, ref):
    a, ref = to_np(a).astype(np.float64), to_np(ref).astype(np.float64)
    scale = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(a - ref).max()) / scale


def assert_close(a, ref, rtol=RTOL, name="", outlier_frac=0.0):
    a, ref = to_np(a).astype(np.float64), to_np(ref).astype(np.float64)
    assert a.shape == ref.shape, f"{name}: shape {a.shape} vs {ref.shape}"
    assert np.isfinite(a).all(), f"{name}: non-finite values"
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = np.abs(a - ref) / scale
    bad = err > rtol
    frac = float(bad.mean())
    _record("tensor", name, {"max_rel": float(err.max()) if err.size else 0.0, "frac_over_rtol": frac,
                              "frac_over_1e-5": float((err > 1e-5).mean()) if err.size else 0.0, "n": int(err.size)},
            {"rtol": rtol, "outlier_frac": outlier_frac})
    assert frac <= outlier_frac, (f"{name}: {bad.sum()} of {bad.size} elements exceed rtol={rtol} "
                                  f"(max rel err {err.max():.3e}, first at {np.argwhere(bad)[0] if bad.any() else None})")


def assert_scalar_close(a, ref, rtol=RTOL, name=""):
    a, ref = float(a), float(ref)
    _record("scalar", name, {"rel": abs(a - ref) / max(abs(ref), 1e-30)}, {"rtol": rtol})
    assert abs(a - ref) <= rtol * max(abs(ref), 1e-30), f"{name}: {a!r} vs {ref!r} (rel {abs(a-ref)/max(abs(ref),1e-30):.3e})"


def assert_mismatch_frac(a, ref, max_frac, name=""):
    """Boolean / index planes compared element-wise: fraction of differing elements (0 = bit-exact)."""
    a, ref = to_np(a), to_np(ref)
    assert a.shape == ref.shape, f"{name}: shape {a.shape} vs {ref.shape}"
    frac = float((a != ref).mean())
    _record("mismatch", name, {"frac": frac, "n": int(a.size)}, {"max_frac": max_frac})
    assert frac <= max_frac, f"{name}: {frac:.3e} of the elements differ (allowed {max_frac:.1e})"
