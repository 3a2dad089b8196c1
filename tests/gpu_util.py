"""Helpers shared by the GPU parity tests."""
import torch

import pcd_b200 as P
from oracle import cases, det
from test_oracle_golden import shapes_of

DEV = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")

# north-star tolerances (BASELINE.json): per-step denoiser output relative L2
TOL_F32 = 1e-4
TOL_BF16 = 2e-2


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def describe(a, b, name=""):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    diff = (a - b).abs()
    idx = int(diff.flatten().argmax())
    return (f"{name}: rel_l2={rel(a, b):.3e} max_abs={float(diff.max()):.3e} at flat {idx} "
            f"(got {float(a.flatten()[idx]):.6f} want {float(b.flatten()[idx]):.6f}) "
            f"nan={int(torch.isnan(a).sum())} ref_std={float(b.std()):.4f} got_std={float(a.float().std()):.4f}")


def build_model(case, dtype):
    """Product model for a forward case with the deterministic case weights."""
    cfg, B, seed, mode = cases.FORWARD_CASES[case]
    sd = det.fill_state_dict(shapes_of(cfg), seed, mode=mode, width=cfg["width"])
    model = P.model_from_config(cases.model_ctor_cfg(cfg), DEV, dtype=dtype)
    missing = model.load_state_dict(sd, strict=True)
    return model.eval(), cfg, sd


def to_dev(kw):
    return {k: v.to(DEV) for k, v in kw.items()}
