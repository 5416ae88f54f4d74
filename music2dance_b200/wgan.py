"""Fused WGAN-GP passes over a CriticNet (losses.py:5-60 + phase3/train.py:204-215).

The gradient penalty is evaluated without autograd:
  1. forward of the critic on the interpolates (ReLU masks kept),
  2. backward-data to both inputs -> g0, g1, per-sample norms, penalty, kappa,
  3. tangent (JVP) pass along v = kappa*g through the same masks,
  4. weight gradients  dW_k = wgrad(delta_k, t_{k-1})  — for a piecewise-linear critic
     d/dtheta <v, grad_x D> = d/dtheta JVP_v(D), whose back-propagated deltas equal the
     first-backward deltas, so no second backward sweep over the data path is needed and
     the identically-zero passes of autograd's double backward never run,
  5. activ='tanh' (tanh.yaml): the code activations c = tanh(z) are the only curved part of
     the critic, so the double backward has exactly one more term, the curvature of tanh:
     <v, J_z^T (s * d(1 - c^2))> = <e_z, dz/dtheta>  with  e_z = -2 c * t_c * s,
     s = dD/dc, t_c = tangent of c.  That is one ordinary backward of each branch from e_z
     (weights AND biases: with tanh the penalty does reach the biases below the codes).
"""
from __future__ import annotations

from . import ops
from .nets import ACT_ID, ACT_RELU, ACT_TANH
from .ops import Mat


def rows(m, b0, b1, per=1):
    """Rows [b0*per, b1*per) of a dense [1, n*per, C] matrix or batches [b0,b1) of [n,L,C]."""
    if m is None:
        return None
    if m.nb == 1:
        v = Mat(m.t, 1, (b1 - b0) * per, m.cols, m.ld, (b1 - b0) * per * m.ld)
        v.ptr = m.ptr + 4 * b0 * per * m.ld
        return v
    return m.batch_slice(b0, b1)


def slice_pose_saves(sv, b0, b1):
    out = {"X": rows(sv["X"], b0, b1), "r0": rows(sv["r0"], b0, b1), "y": rows(sv["y"], b0, b1),
           "code": rows(sv["code"], b0, b1),
           "blk": [tuple(rows(t, b0, b1) for t in blk) for blk in sv["blk"]]}
    return out


def critic_forward(D, X, audio, nS, nA, tag, groups=1):
    """Pose branch on X [nS,T,O], audio branch on audio [nA,A] (shared by `groups`
    row groups of nA samples each: Q13 de-duplication), fusion MLP on nS rows."""
    wk = D.wk
    sa = wk.mat(f"{tag}:sa", 1, nS, D.F)
    sva = None
    if not D.ablated:
        with D.fork():                                   # audio branch: side stream
            sva = D.audio_fwd(audio, nA, tag)
            for g in range(groups):
                ops.copy2d(sva["code"], rows(sa, g * nA, (g + 1) * nA).cols_slice(D.code, D.F))
    svp = D.pose_fwd(X, nS, tag)
    ops.copy2d(svp["code"], sa.cols_slice(0, D.code))
    D.join()
    u, d = D.fusion_fwd(sa, nS, tag)
    return dict(svp=svp, sva=sva, sa=sa, u=u, d=d)


def gradient_penalty_pass(D, fw, B, tag, scale, beta, gp_out, k0, k1, weight_grads=True, before_wgrads=None):
    """GP on the first B rows of forward state `fw`.  Writes gp to gp_out[0] and, if
    weight_grads, accumulates  scale * dGP/dW  into the critic's gradient buffers
    (gw = beta*gw + ...; biases receive nothing, Q5)."""
    wk, T, O, code = D.wk, D.T, D.O, D.code
    if D.act not in (ACT_ID, ACT_RELU, ACT_TANH) and weight_grads:
        raise NotImplementedError("gradient-penalty weight gradients: activ must be 'id', 'relu' or 'tanh'")
    svp = slice_pose_saves(fw["svp"], 0, B)
    sva = fw["sva"]
    u = rows(fw["u"], 0, B)
    ones = wk.vec("ones", B)
    ops.fill(ones, B, 1.0)
    dh, dsa = D.fusion_bwd(Mat(ones, 1, B, 1), u, B, tag)
    g1 = ss1 = None
    if not D.ablated:
        A = D.cfg["audio_length"]
        ss1 = wk.acc_slot(B)
        g1 = wk.vec(f"{tag}:g1", B * A)
        with D.fork():
            d_a = wk.mat(f"{tag}:d_a", 1, B, code)
            ops.copy2d(dsa.cols_slice(code, D.F), d_a)
            D.audio_bwd(sva, d_a, B, tag, wgrads=False, dX=g1)
            ops.rows_sumsq(g1, B, A, ss1)
    d_s = wk.mat(f"{tag}:d_s", 1, B, code)
    ops.copy2d(dsa.cols_slice(0, code), d_s)
    g0 = wk.mat(f"{tag}:g0", B, T, O)
    D.pose_bwd(svp, d_s, B, tag, wgrads=False, dX=g0)
    ss0 = wk.acc_slot(B)
    ops.rows_sumsq(g0, B, T * O, ss0)
    D.join()
    ops.gp_finalize(ss0, ss1, B, gp_out, k0, k1)
    out = dict(g0=g0, g1=g1)
    if not weight_grads:
        return out
    # tangent pass along v = kappa * g
    t_sa = wk.mat(f"{tag}:t_sa", 1, B, D.F)
    tva = None
    if not D.ablated:
        with D.fork():
            ops.scale_rows(g1, k1, g1, B, A)
            t_a = wk.mat(f"{tag}:t_a", 1, B, code)
            tva = D.audio_tangent(sva, g1, B, tag, t_a)
            ops.copy2d(t_a, t_sa.cols_slice(code, D.F))
    ops.scale_rows(g0, k0, g0, B, T * O)
    t_s = wk.mat(f"{tag}:t_s", 1, B, code)
    tvp = D.pose_tangent(svp, g0, B, tag, t_s)
    ops.copy2d(t_s, t_sa.cols_slice(0, code))
    D.join()
    t_h = wk.mat(f"{tag}:t_h", 1, B, 128)
    D.fc1.fwd(t_sa, t_h, bias=False, ws=wk.scratch, mask=u, mask_mode=ACT_RELU)
    # weight gradients: wgrad(first-backward delta, tangent activation)
    if before_wgrads is not None:
        before_wgrads()                  # e.g. wait for a concurrent chain that writes the same gradient buffers first
    if not D.ablated:
        with D.fork():
            D.audio_wgrads(sva["delta"], tva["X"], tva["q"], scale, beta, bias=False)
    ops.colsum(t_h, D.fc2.gw, wk.acc_slot(128), scale=scale, beta=beta)
    D.fc1.wgrad(dh, t_sa, wk.scratch, scale=scale, beta=beta, bias=False)
    D.pose_wgrads(svp["delta"], g0, tvp, scale, beta, bias=False)
    D.join()
    if D.act == ACT_TANH:
        # curvature of the code activations: ordinary backward of both branches from
        # e_z = -2 c * t_c * s (dsa still holds s = dD/dc: the first backward worked on copies)
        if not D.ablated:
            with D.fork():
                e_a = wk.mat(f"{tag}:e2_a", 1, B, code)
                ops.mul3(dsa.cols_slice(code, D.F), t_a, sva["code"], e_a, alpha=-2.0)
                D.audio_bwd(dict(sva), e_a, B, f"{tag}2", scale=scale, beta=1.0, wgrads=True, bbeta=beta,
                            pre_act=True)
        e_s = wk.mat(f"{tag}:e2_s", 1, B, code)
        ops.mul3(dsa.cols_slice(0, code), t_s, svp["code"], e_s, alpha=-2.0)
        D.pose_bwd(dict(svp), e_s, B, f"{tag}2", scale=scale, beta=1.0, wgrads=True, bbeta=beta, pre_act=True)
        D.join()
    return out


def wasserstein_backward(D, fw, r0, nR, signs, B, tag, beta, dX_rows=None, dX=None, param_grads=True):
    """Backward of  sum_g signs[g] * mean(D(rows of group g))  over `len(signs)` groups of B
    rows starting at row r0 of forward state `fw`.  Weight grads: gw = beta*gw + dW, biases
    overwritten.  If dX is given (Mat [B,T,O]) the pose-input gradient of group
    `dX_rows` is written there (generator update)."""
    wk, code = D.wk, D.code
    n = nR
    dd = wk.vec(f"{tag}:dd", n)
    for g, s in enumerate(signs):
        ops.fill(dd[g * B:(g + 1) * B], B, s / B)
    ddm = Mat(dd, 1, n, 1)
    u = rows(fw["u"], r0, r0 + n)
    sa = rows(fw["sa"], r0, r0 + n)
    if param_grads:
        D.fc2.wgrad(ddm, u, wk.scratch, beta=beta, bbeta=0.0, acc=wk.acc_slot(1))
    dh, dsa = D.fusion_bwd(ddm, u, n, tag)
    if param_grads:
        D.fc1.wgrad(dh, sa, wk.scratch, beta=beta, bbeta=0.0, acc=wk.acc_slot(128))
    if not D.ablated and param_grads:
        with D.fork():
            d_a = wk.mat(f"{tag}:d_a", 1, B, code)
            for g in range(len(signs)):
                ops.copy2d(rows(dsa, g * B, (g + 1) * B).cols_slice(code, D.F), d_a, accumulate=(g > 0))
            D.audio_bwd(fw["sva"], d_a, B, tag, scale=1.0, beta=beta, wgrads=True, bbeta=0.0)
    d_s = wk.mat(f"{tag}:d_s", 1, n, code)
    ops.copy2d(dsa.cols_slice(0, code), d_s)
    svp = slice_pose_saves(fw["svp"], r0, r0 + n)
    if param_grads:
        D.pose_bwd(svp, d_s, n, tag, scale=1.0, beta=beta, wgrads=True, bbeta=0.0)
    else:
        # generator update: only the input gradient of one group is needed
        g = dX_rows
        D.pose_bwd(slice_pose_saves(svp, g * B, (g + 1) * B), rows(d_s, g * B, (g + 1) * B), B, tag,
                   wgrads=False, dX=dX)
    D.join()
