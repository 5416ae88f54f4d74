"""Shared helpers for the parity tests (golden fixtures + tolerances)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = {"default": {}, "wavegan": {"enc_type": "wavegan"}, "unet": {"enc_type": "unet"},
            "ablated": {"ablated": True}, "tanh": {"activ": "tanh"}}
B_GOLD, ALPHA_SEED, DATA_SEED = 2, 77, 1234

# tolerance stated by BASELINE.json's north_star: 1e-3 relative on losses and GP terms.
# fp32 paths (oracle, SIMT kernels) are held to a much tighter bound.
TOL_NORTH_STAR = 1e-3
TOL_FP32 = 2e-4
# gradients of the widest reductions (first encoder conv: K = B*T*L_out ~ 1e6 terms) move by a
# few 1e-4 between two fp32 summation orders of the SAME reference code (thread count), so
# gradient digests are held to the north-star figure rather than the scalar one.
TOL_GRAD = 1e-3
# The generator loss contains L1 = mean|real - fake| whose gradient is sign(real - fake)/n:
# DISCONTINUOUS.  An element with |real - fake| below the fp32 noise of `fake` flips sign
# between two summation orders of the same reference code and moves every generator
# gradient by ~2/n of its scale (n = B*T*69 = 16560 at B = 2).  End-to-end generator-gradient
# digests therefore allow a handful of flips; the smooth-upstream component tests
# (test_generator_backward_smooth) hold the same kernels to TOL_GRAD.
TOL_GEN_GRAD_E2E = 5e-3


def load_golden(name):
    return np.load(os.path.join(GOLD, f"phase3_{name}.npz"))


def digest_check(t, gold, prefix, tol, what, abs_floor=0.0):
    """Compare a tensor against a stored (sum, l2, maxabs, strided samples) digest."""
    f = t.detach().double().cpu().reshape(-1)
    n = f.numel()
    step = max(1, n // 64)
    samp = f[::step][:64]
    gs = torch.from_numpy(gold[prefix + "/samples"]).double()
    if prefix.startswith("window") or gs.numel() != samp.numel():
        step = max(1, n // gs.numel())
        samp = f[::step][:gs.numel()]
    scale = max(float(gold[prefix + "/maxabs"]), abs_floor, 1e-30)
    e_s = float((samp - gs).abs().max()) / scale
    assert e_s < tol, f"{what}: samples deviate {e_s:.3e} (rel. to max |x| = {scale:.3e})"
    l2 = float(gold[prefix + "/l2"])
    e_l2 = abs(float(f.norm()) - l2) / max(l2, abs_floor * n ** 0.5, 1e-30)
    assert e_l2 < tol, f"{what}: l2 deviates {e_l2:.3e}"
    e_sum = abs(float(f.sum()) - float(gold[prefix + "/sum"])) / max(l2 * n ** 0.5, 1e-30)
    assert e_sum < tol, f"{what}: sum deviates {e_sum:.3e}"


def scalar_check(v, ref, tol, what):
    e = abs(float(v) - float(ref)) / max(abs(float(ref)), 1e-2)
    assert e < tol, f"{what}: {float(v):.7f} vs {float(ref):.7f} (rel {e:.3e})"
