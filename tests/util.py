"""Shared helpers of the test suite (golden loading, tolerance rule, module construction)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARAM_KEYS = ["scalar", "conv_down.bias", "conv_down.lin.weight", "conv_up.bias", "conv_up.lin.weight"]

# north_star: "within rtol 1e-5 in fp32".  The r-wide reordering (Ahat Z) Wu^T changes the fp32
# summation order, so elements that cancel to ~0 cannot meet a pure relative bound
# ("Hard parts"): the absolute floor is 1e-5 of the tensor's largest magnitude.
RTOL = 1e-5
ATOL_SCALE = 1e-5


def golden_names(prefix: str = ""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {k: z[k] for k in z.files}
    out["cfg"] = {k[4:]: out[k].item() for k in list(out) if k.startswith("cfg.")}
    return out


def ctor_kwargs(cfg: dict) -> dict:
    return {k: v for k, v in cfg.items() if k != "three_d"}


def assert_close(ours, ref, what: str, rtol: float = RTOL, atol_scale: float = ATOL_SCALE, max_outlier_frac: float = 0.0,
                 noise_floor: float = 0.0):
    """|ours - ref| <= atol_scale * max|ref| + rtol * |ref| (+ noise_floor, for quantities that are
    mathematically zero so that the reference itself holds only rounding noise)."""
    ours = torch.as_tensor(ours).detach().double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).detach().double().cpu().reshape(-1)
    assert ours.shape == ref.shape, f"{what}: shape {tuple(ours.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(ours).all(), f"{what}: non-finite values"
    if ref.numel() == 0:
        return
    scale = ref.abs().max().item()
    tol = atol_scale * scale + rtol * ref.abs() + noise_floor
    bad = (ours - ref).abs() > tol
    nbad = int(bad.sum())
    allowed = int(max_outlier_frac * ref.numel())
    if nbad > allowed:
        err = ((ours - ref).abs() / (scale + 1e-300)).max().item()
        raise AssertionError(f"{what}: {nbad}/{ref.numel()} elements outside rtol={rtol}, atol={atol_scale}*max|ref| "
                             f"(allowed {allowed}); max error / max|ref| = {err:.3e}")


def load_module_params(module, params: dict) -> None:
    sd = module.state_dict()
    with torch.no_grad():
        for k, v in params.items():
            if k in sd:
                sd[k].copy_(torch.as_tensor(v))
