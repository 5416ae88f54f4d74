"""Shared helpers for the parity tests (golden fixtures + tolerances)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = {"default": {}, "wavegan": {"enc_type": "wavegan"}, "unet": {"enc_type": "unet"},
            "ablated": {"ablated": True}, "tanh": {"activ": "tanh"},
            # phase3/configs/tv.yaml (eta = 50), noise_enhanced.yaml (noise 100 -> audio GRU hidden 150, noise GRU 100),
            # activ = 'relu' (default.py:73-74,100-101,254)
            "tv": {"eta": 50.0}, "noise_enhanced": {"noise_size": 100}, "relu": {"activ": "relu"}}
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
# (test_generator_backward_smooth) hold the same kernels to TOL_GRAD.  Set equal to TOL_KINK_L2
# below: on B200 one flipped decoder ReLU unit moved the first encoder BatchNorm's weight-gradient
# digest by 5.3e-3 of max|g| against the reference fixture (default/perturbed), every other
# generator gradient of that run staying inside 5e-3.
TOL_GEN_GRAD_E2E = 1e-2
# Measured on B200 (tools/diag_gen.py, fp64 oracle as ground truth): without a kink flip every
# generator gradient of the CUDA path is within 1-2e-6 of fp64 — the same as the fp32 reference.
# One flipped ReLU unit in a layer of U units moves ALL upstream gradients by ~1/sqrt(U) in l2
# (observed: one unit of decoder.blocks.1, U = 360*256 -> 4e-3 everywhere upstream, 5e-2 on the
# flipped channel's own entries), and with ~4e6 ReLU units per generator pass about one unit per
# test point lies inside fp32 rounding noise of zero.  In-test comparisons against the oracle
# therefore bound the l2 error by TOL_KINK_L2 and single entries by TOL_KINK_MAX.
TOL_KINK_L2 = 1e-2
TOL_KINK_MAX = 6e-2
# Chained iterations: Adam's first steps move every weight by ~lr*sign(g), so a flip (or the sign
# of a noise-level gradient) changes the state the next iteration starts from; scalars after the
# first optimiser step are compared at TOL_CHAINED instead of the single-iteration 1e-3.
# Measured on B200 in BOTH arithmetic modes (fp32 CUDA cores / 3xTF32 tensor cores): after two Adam
# steps the small ablated-critic penalty gp = 0.0116 (||g|| ~ 0.9, so d gp / d||g|| amplifies 20x)
# differs from the CPU oracle by 1.1-1.4 %; everything else stays below 1 %.
TOL_CHAINED = 2e-2
# Parameter drift after n Adam steps, in units of lr*n: Adam's first updates are lr*sign-like, so an
# entry whose gradient is at noise level (|g| below the fp32 summation noise) moves by +-lr per step
# in a direction that no two fp32 evaluations agree on.  Measured: 5-6 % of the first encoder
# convolution's 8000 weights are in that regime (mean |dW| = 0.10-0.12 lr*n in both arithmetic modes).
TOL_DRIFT_LR = 0.25
# BatchNorm running statistics follow the drifting weights (momentum 0.1 per forward); compared
# relative to max(max|stat|, 0.1).  Measured: 1.1e-4 .. 5.8e-4 absolute.
TOL_DRIFT_BUF = 1e-2
# Bias gradients of the critic's convolutions: sum over ALL positions of (delta_fake - delta_real),
# two nearly equal contributions (same audio, same weights), i.e. a cancellation-dominated reduction
# whose conditioning amplifies the summation noise of the deltas ~100x.  fp32 CUDA-core kernels land at
# 3e-4 of max|g|, the 3xTF32 tensor-core split (2-4x the fp32 summation noise) at 1.7e-3 .. 7.6e-3.
TOL_GRAD_BIAS = 1e-2


def load_golden(name):
    return np.load(os.path.join(GOLD, f"phase3_{name}.npz"))


def digest_check(t, gold, prefix, tol, what, abs_floor=0.0, kinks=False):
    """Compare a tensor against a stored (sum, l2, maxabs, strided samples) digest.

    kinks=False: every sample, the l2 norm and the sum within `tol` (relative to max|x|).
    kinks=True (gradients that pass through ReLU / LeakyReLU / MaxPool / L1-sign kinks): the
    gradient is a DISCONTINUOUS function of the inputs, so two fp32 evaluations of the same
    network (the reference on 8 vs 16 threads, the oracle vs the reference, CUDA vs CPU) differ
    by whole units whose pre-activation lies within rounding noise of zero: isolated entries
    move by percents while everything else agrees to ~1e-4.  The check is therefore statistical:
    90 % of the samples within `tol`, every sample within 10*tol, l2 and sum within 4*tol."""
    f = t.detach().double().cpu().reshape(-1)
    n = f.numel()
    step = max(1, n // 64)
    samp = f[::step][:64]
    gs = torch.from_numpy(gold[prefix + "/samples"]).double()
    if prefix.startswith("window") or gs.numel() != samp.numel():
        step = max(1, n // gs.numel())
        samp = f[::step][:gs.numel()]
    scale = max(float(gold[prefix + "/maxabs"]), abs_floor, 1e-30)
    dev = (samp - gs).abs() / scale
    if kinks:
        q90 = float(dev.sort().values[max(0, int(0.9 * dev.numel()) - 1)])
        assert q90 < tol, f"{what}: 90 % quantile of sample deviations {q90:.3e} (rel. to max |x| = {scale:.3e})"
        assert float(dev.max()) < 10 * tol, f"{what}: worst sample deviates {float(dev.max()):.3e}"
        tol = 4 * tol
    else:
        e_s = float(dev.max())
        assert e_s < tol, f"{what}: samples deviate {e_s:.3e} (rel. to max |x| = {scale:.3e})"
    l2 = float(gold[prefix + "/l2"])
    e_l2 = abs(float(f.norm()) - l2) / max(l2, abs_floor * n ** 0.5, 1e-30)
    assert e_l2 < tol, f"{what}: l2 deviates {e_l2:.3e}"
    e_sum = abs(float(f.sum()) - float(gold[prefix + "/sum"])) / max(l2 * n ** 0.5, 1e-30)
    assert e_sum < tol, f"{what}: sum deviates {e_sum:.3e}"


def scalar_check(v, ref, tol, what):
    e = abs(float(v) - float(ref)) / max(abs(float(ref)), 1e-2)
    assert e < tol, f"{what}: {float(v):.7f} vs {float(ref):.7f} (rel {e:.3e})"
