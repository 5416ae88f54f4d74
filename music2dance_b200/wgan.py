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

import torch

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
            ops.copy2d_batch([(sva["code"], rows(sa, g * nA, (g + 1) * nA).cols_slice(D.code, D.F), False)
                              for g in range(groups)])
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


def critic_forward_fused(D, X3, audio, B, tag, before_pose=None):
    """critic_forward for the fused backward below: pose branch on the 3B rows [interpolates; real; fake], audio
    branch once on B samples with duplicated activation buffers (CriticNet.audio_fwd(dup=True)).
    before_pose: called after the audio branch has been forked and before the pose branch starts — the audio branch
    depends on neither the generated poses nor X3, so the caller waits for the generator forward / builds X3 there."""
    wk = D.wk
    n3 = 3 * B
    sa = wk.mat(f"{tag}:sa", 1, n3, D.F)
    sva = None
    if not D.ablated:
        with D.fork():
            sva = D.audio_fwd(audio, B, tag, dup=True)
            ops.copy2d_batch([(sva["code"], rows(sa, g * B, (g + 1) * B).cols_slice(D.code, D.F), False)
                              for g in range(3)])
    if before_pose is not None:
        before_pose()
    svp = D.pose_fwd(X3, n3, tag)
    ops.copy2d(svp["code"], sa.cols_slice(0, D.code))
    D.join()
    # the upstream of the scores is a constant of the step, so the fusion MLP's backward-data rides in its forward launch
    ddm = score_upstream(D, B)
    u, d, dh, dsa = D.fusion_fwd(sa, n3, "f", dd=ddm)
    ops.mark(f"{tag}:fwd_end")
    return dict(svp=svp, sva=sva, sa=sa, u=u, d=d, ddm=ddm, dh=dh, dsa=dsa)


def score_upstream(D, B):
    """d(err_fake - err_real + gamma*gp) / d(scores) over the rows [interpolates; real; fake] up to the penalty's
    kappa (applied later): 1, -1/B, +1/B — a constant vector, written once."""
    wk = D.wk
    n3 = 3 * B
    dd = wk.vec("f:dd", n3)
    consts = D.__dict__.setdefault("_consts", set())
    if ("f", "dd", B) not in consts:
        ops.fill(dd[:B], B, 1.0)
        ops.fill(dd[B:2 * B], B, -1.0 / B)
        ops.fill(dd[2 * B:], B, 1.0 / B)
        consts.add(("f", "dd", B))
    return Mat(dd, 1, n3, 1)


def critic_backward_fused(D, fw, B, audio, gamma, gp_out, k0, k1, tag="f", aud2=None, early_reduce=None):
    """Critic gradients of  err_fake - err_real + gamma * gp  (phase3/train.py:204-215) in ONE backward sweep.

    Rows of the forward state: [0,B) interpolates, [B,2B) real, [2B,3B) fake.  The Wasserstein terms and the
    penalty are back-propagated together — pose branch: one backward-data pass over all 3B rows (upstream 1 on
    the interpolates, -1/B / +1/B on real / fake); audio branch (evaluated once on B samples): one pass over 2B
    batch entries [Wasserstein upstream; penalty upstream] through the same ReLU masks (second copy of the
    activations).  After the per-sample norms, v = gamma * kappa * g goes through the tangent pass IN PLACE: the
    tangent activations overwrite the interpolates' forward activations (pose) / the mask copies (audio), so that
    afterwards every layer has ONE stacked input operand [tangent; forward] matching its stacked deltas
    [penalty; Wasserstein], and a single weight-gradient GEMM per layer produces
        dW = sum delta_w (x) x  +  gamma * sum delta_gp (x) t.
    Against the two separate chains (wasserstein_backward + gradient_penalty_pass) this halves the backward-data
    and weight-gradient launches of an iteration.  Biases: Wasserstein rows only (Q5).  Piecewise-linear critic
    with identity code activation only (activ='id'); other settings use the two-chain path."""
    assert D.act == ACT_ID
    wk, T, O, code = D.wk, D.T, D.O, D.code
    n3 = 3 * B
    A1 = lambda c: wk.acc_slot(c)
    ddm = fw["ddm"]                                               # upstream of the scores: a constant (score_upstream)
    u, sa = fw["u"], fw["sa"]
    dh, dsa = fw["dh"], fw["dsa"]                                 # fusion MLP backward-data: done by the forward launch
    ops.mark(f"{tag}:fusion_bwd")
    g1 = ss1 = None
    sva = fw["sva"]
    if not D.ablated:
        Alen = D.cfg["audio_length"]
        ss1 = wk.acc_slot(B)
        g1 = aud2 if aud2 is not None else wk.mat(f"{tag}:aud2", 2 * B, Alen, 1)   # [audio; v1]: stacked input of l1
        with D.fork():
            d_a2 = wk.mat(f"{tag}:d_a2", 1, 2 * B, code)
            ops.copy2d_batch([(rows(dsa, B, 2 * B).cols_slice(code, D.F), rows(d_a2, 0, B), False),
                              (rows(dsa, 2 * B, n3).cols_slice(code, D.F), rows(d_a2, 0, B), True),
                              (rows(dsa, 0, B).cols_slice(code, D.F), rows(d_a2, B, 2 * B), False)])
            sv2 = {"q": sva["q2"], "code": None, "X": None}
            dla = D.audio_bwd(sv2, d_a2, 2 * B, tag, wgrads=False, dX=None)
            gv = g1.batch_slice(B, 2 * B)
            D.l1_dgrad(dla[0].batch_slice(B, 2 * B), gv, B)
            ops.mark(f"{tag}:aud_bwd_l1")
            ops.rows_sumsq(gv, B, Alen, ss1)
            if aud2 is None:
                ops.copy2d(Mat.of(audio.reshape(-1), 1, B, Alen) if not isinstance(audio, Mat) else audio.flat_rows(),
                           Mat(g1.t, 1, B, Alen, Alen))
    d_s3 = wk.mat(f"{tag}:d_s3", 1, n3, code)
    ops.copy2d(dsa.cols_slice(0, code), d_s3)
    svp = fw["svp"]
    dlp = D.pose_bwd(svp, d_s3, n3, tag, wgrads=False, dX=None)
    D.fc2.wgrad(rows(ddm, B, n3), rows(u, B, n3), wk.scratch, beta=0.0, bias=False)
    D.fc1.wgrad(rows(dh, B, n3), rows(sa, B, n3), wk.scratch, beta=0.0, bias=False)
    X3 = svp["X"]
    g0 = wk.mat(f"{tag}:g0", B, T, O)
    D.s_conv1.dgrad(rows(dlp["conv1"], 0, B), g0, ws=wk.scratch)
    ss0 = wk.acc_slot(B)
    ops.rows_sumsq(g0, B, T * O, ss0)
    D.join()
    # kappa scaled by gamma in the same launch: v = gamma * kappa * g is written where the layer-1 weight gradients
    # read their input
    kg = wk.vec(f"{tag}:kg", 2 * B)
    ops.gp_finalize(ss0, ss1, B, gp_out, kg[:B], kg[B:], kscale=gamma)
    ops.mark(f"{tag}:gp")
    t_sa = wk.mat(f"{tag}:t_sa", 1, B, D.F)
    main = torch.cuda.current_stream(D.dev)
    par = D.par
    # The weight-gradient GEMMs are leaves of the dependency graph: each needs its deltas (ready) and ONE tangent
    # activation.  They run on their own streams (s_wa: audio, s_w: pose), every launch waiting only for the
    # tangent layer that produces its input, so they fill the gaps of the (latency-bound) tangent chains.
    s_ta = D.s_aud if par else main
    s_wa = D.s_wa if par else main
    s_wp = D.s_w if par else main

    def after(ev_stream):
        ev = torch.cuda.Event()
        ev.record(ev_stream)
        return ev

    if not D.ablated:
        if par:
            s_ta.wait_stream(main)
            s_wa.wait_stream(main)
        with torch.cuda.stream(s_ta):
            gv = g1.batch_slice(B, 2 * B)
            ops.scale_rows(gv, kg[B:], gv, B, Alen)
            evs = [after(s_ta)]
            x = gv
            for i, l in enumerate(D.a_layers):                    # tangent pass in place over the mask copies
                t = sva["q2"][i].batch_slice(B, 2 * B)
                l.fwd(x, t, bias=False, ws=wk.scratch, mask=t, mask_mode=ACT_RELU)
                evs.append(after(s_ta))
                ops.mark(f"{tag}:aud_tan_l{i + 1}")
                x = t
            t_a = wk.mat(f"{tag}:t_a", 1, B, code)
            D.a_l6.fwd(x, t_a.as_rows(B, 1), bias=False, ws=wk.scratch)
            ops.copy2d(t_a, t_sa.cols_slice(code, D.F))
            ops.mark(f"{tag}:aud_tan_l6")
        with torch.cuda.stream(s_wa):
            # one weight-gradient GEMM per audio layer over the 2B stacked entries; biases from the Wasserstein half
            x = g1
            ws_a = wk.scratch                                      # per-stream scratch (Workspace.scratch)
            ev_wa_early = None
            for i, l in enumerate(D.a_layers):
                if par:
                    s_wa.wait_event(evs[i])
                l.wgrad(dla[i], x, ws_a, scale=1.0, beta=0.0, bias=False)
                ops.mark(f"{tag}:aud_wg_l{i + 1}")
                x = sva["q2"][i]
                if i == len(D.a_layers) - 2 and early_reduce is not None and par:
                    ev_wa_early = after(s_wa)                     # audio_d.l1 .. l4 weight gradients done
            if par:
                s_wa.wait_event(evs[len(D.a_layers)])
            D.a_l6.wgrad(d_a2.as_rows(2 * B, 1), x, ws_a, scale=1.0, beta=0.0, bias=False)
            ops.mark(f"{tag}:aud_wg_l6")
    # pose branch: tangent in place on the current stream, weight gradients over all 3B entries on s_w
    Xi = rows(X3, 0, B)
    ops.scale_rows(g0, kg[:B], Xi, B, T * O)                       # interpolates are no longer needed: X3[0:B] = v0
    if par:
        s_wp.wait_stream(main)
    sp = slice_pose_saves(svp, 0, B)
    pending = []                                                   # (conv, deltas, input operand, event)

    def wg(conv, dl, x_in, ev):
        pending.append((conv, dl, x_in, ev))

    wg(D.s_conv1, dlp["conv1"], X3, after(main))
    t0 = sp["r0"]
    D.s_conv1.fwd(Xi, t0, bias=False, ws=wk.scratch, mask=t0, mask_mode=ACT_RELU)
    ev = after(main)
    x = t0
    for b, (c1, c2) in enumerate(D.s_blocks):
        x_fw, r1_fw, r2_fw, y_fw = svp["blk"][b]
        _, r1, r2, y = sp["blk"][b]
        wg(c1, dlp[f"b{b}c1"], x_fw, ev)
        c1.fwd(x, r1, bias=False, ws=wk.scratch, mask=r1, mask_mode=ACT_RELU)
        ev = after(main)
        wg(c2, dlp[f"b{b}c2"], r1_fw, ev)
        c2.fwd(r1, y, bias=False, ws=wk.scratch, mask=r2, mask_mode=ACT_RELU, add=x)
        ev = after(main)
        x = y
    wg(D.s_fconv, d_s3.as_rows(n3, 1), svp["y"], ev)
    t_s = wk.mat(f"{tag}:t_s", 1, B, code)
    D.s_fconv.fwd(x, t_s.as_rows(B, 1), bias=False, ws=wk.scratch)
    ops.copy2d(t_s, t_sa.cols_slice(0, code))
    ops.mark(f"{tag}:pose_tan_end")
    with torch.cuda.stream(s_wp):
        ws_p = wk.scratch
        # every bias gradient of the iteration (Wasserstein rows only, Q5) in one batched column-sum launch: the deltas
        # of both branches are complete (the main stream joined the audio branch before the penalty was finalised)
        tabs = D.__dict__.setdefault("_bias_tabs", {})
        key = (tag, B, tuple(dl.ptr for _, dl, _, _ in pending))
        if key not in tabs:
            ent = [(rows(ddm, B, n3), D.fc2.gb, 1.0, 0.0), (rows(dh, B, n3), D.fc1.gb, 1.0, 0.0)]
            for conv, dl, _, _ in pending:
                ent.append((rows(d_s3, B, n3) if conv is D.s_fconv else rows(dl, B, n3).flat_rows(), conv.gb, 1.0, 0.0))
            if not D.ablated:
                for i, l in enumerate(D.a_layers):
                    ent.append((dla[i].batch_slice(0, B).flat_rows(), l.gb, 1.0, 0.0))
                ent.append((rows(d_a2, 0, B), D.a_l6.gb, 1.0, 0.0))
            tabs[key] = ops.colsum_table(ent, None, D.dev)
        ops.colsum_batch(tabs[key])
        for conv, dl, x_in, ev in pending:
            if par:
                s_wp.wait_event(ev)
            conv.wgrad(dl, x_in, ws_p, scale=1.0, beta=0.0, bias=False)
        ops.mark(f"{tag}:pose_wg_end")
    early_done = False
    if early_reduce is not None and par and not D.ablated and ev_wa_early is not None:
        s_comm = D.comm_stream()
        s_comm.wait_stream(s_wp)
        s_comm.wait_event(ev_wa_early)
        with torch.cuda.stream(s_comm):
            early_reduce()
            ops.mark(f"{tag}:early_reduce")
        early_done = True
    if par and not D.ablated:
        main.wait_stream(s_ta)
    # fusion MLP: penalty part through the tangent of the codes
    t_h = wk.mat(f"{tag}:t_h", 1, B, 128)
    D.fc1.fwd(t_sa, t_h, bias=False, ws=wk.scratch, mask=rows(u, 0, B), mask_mode=ACT_RELU)
    ops.colsum(t_h, D.fc2.gw, A1(128), scale=1.0, beta=1.0)
    D.fc1.wgrad(rows(dh, 0, B), t_sa, wk.scratch, scale=1.0, beta=1.0, bias=False)
    if par:
        if not D.ablated:
            main.wait_stream(s_wa)
        main.wait_stream(s_wp)
        if early_done:
            main.wait_stream(s_comm)
    ops.mark(f"{tag}:bwd_end")
    return early_done


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
