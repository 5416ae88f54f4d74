"""Drop-in for the hot-path functions of the reference's ``losses.py``:
``gradient_penalty`` (losses.py:5-60, the two sequence branches phase3 uses) and
``tv_loss`` (losses.py:76-82).  Same signatures; results participate in autograd."""
from __future__ import annotations

import torch

from . import ops
from .ops import Mat


def _channels_last(t, B, C, L, wk, name):
    """(B,C,L) tensor -> dense channels-last Mat [B,L,C] (copy; handles permuted views, Q8)."""
    X = wk.mat(name, B, L, C)
    tt = t.detach().transpose(1, 2)
    if tt.is_contiguous() and tt.dtype == torch.float32:
        ops.copy2d(Mat.of(tt, 1, B * L, C), X.flat_rows())
    else:
        ops.transpose_bcl(t.detach().contiguous().float(), X, B, C, L)
    return X


class _GradPenaltyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, critic, real, fake, audio, alpha, *params):
        from .wgan import critic_forward, gradient_penalty_pass
        eng = critic._engine()
        eng.ensure_packed()
        D = eng.net
        wk = D.wk
        B = real.shape[0]
        with torch.cuda.device(real.device):
            wk.acc_reset()
            R = _channels_last(real.view(B, D.O, -1), B, D.O, D.T, wk, "gp:real")
            Fk = _channels_last(fake.view(B, D.O, -1), B, D.O, D.T, wk, "gp:fake")
            X = wk.mat("gp:xi", B, D.T, D.O)
            ops.interp(R, Fk, alpha.contiguous().float(), X, B, D.T * D.O)
            aud = None if D.ablated else audio.detach().contiguous().float().view(B, -1)
            fw = critic_forward(D, X, aud, B, B, "gp")
            gp = torch.zeros(1, device=real.device)
            k0, k1 = wk.vec("gp:k0", B), wk.vec("gp:k1", B)
            gradient_penalty_pass(D, fw, B, "gp", 1.0, 0.0, gp, k0, k1, weight_grads=True)
            ctx.grads = [None if g is None else g for g in eng.grads_in_param_order()]
            # GP gives no gradient to biases (Q5): the bias slots were not written by this pass —
            # except below a tanh code activation, whose curvature term reaches the branch biases
            names = eng.fp.names
            tanh = critic.activ == "tanh" if hasattr(critic, "activ") else False
            live_bias = lambda n: tanh and not n.startswith("fc")
            ctx.grads = [g if (g is not None and (not n.endswith(".bias") or live_bias(n))) else None
                         for n, g in zip(names, ctx.grads)]
            return gp.view(())

    @staticmethod
    def backward(ctx, dgp):
        return (None, None, None, None, None, *[None if g is None else g * dgp for g in ctx.grads])


def gradient_penalty(critic, bsize, real, fake, audio=None, is_seq=False, is_cond=False, lp=False, device=None):
    """WGAN-GP penalty mean((||grad_x D(x_hat)|| - 1)^2) [+ the same for the audio input].

    `alpha` is drawn with ``torch.rand(bsize, 1)`` on the CPU generator exactly like the
    reference (Q3), so seeding reproduces the reference's interpolates.  Only the branch
    phase3/train.py uses is implemented: is_seq=True, is_cond=False, lp=False."""
    if not is_seq or is_cond or lp:
        raise NotImplementedError("music2dance_b200.losses.gradient_penalty implements the phase3 call "
                                  "(is_seq=True, is_cond=False, lp=False); phase1/phase2 variants are out of scope")
    if not real.is_cuda:
        raise RuntimeError("gradient_penalty needs CUDA tensors (no CPU fallback)")
    alpha = torch.rand(bsize, 1).to(real.device).view(-1)
    return _GradPenaltyFn.apply(critic, real, fake, audio, alpha, *critic.parameters())


class _TVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq):
        B, C, T = seq.shape
        with torch.cuda.device(seq.device):
            st = seq.detach().transpose(1, 2)
            x = st if (st.is_contiguous() and st.dtype == torch.float32) else st.contiguous().float()
            acc = torch.zeros(2, dtype=torch.float64, device=seq.device)
            d = torch.empty(B, T, C, device=seq.device)
            ops.pose_losses(x, x, d, B, T, C, 0.0, 1.0, False, acc)
            ctx.d = d
            return (acc[1] / (B * (T - 1) * C)).float()

    @staticmethod
    def backward(ctx, g):
        return (ctx.d * g).transpose(1, 2)


def jerkiness(sequence):
    """Smoothness metric of phase3/test.py:85-100 (losses.py:85-89): squared third temporal difference, summed
    over the coordinates, averaged over batch and time, for (B, C, T) input.  No gradient (evaluation only)."""
    if not sequence.is_cuda:
        raise RuntimeError("jerkiness needs CUDA tensors (no CPU fallback)")
    B, C, T = sequence.shape
    with torch.cuda.device(sequence.device):
        x = sequence.detach().transpose(1, 2).contiguous().float()
        acc = torch.zeros(1, dtype=torch.float64, device=sequence.device)
        ops.jerkiness(x, B, T, C, acc)
        return (acc[0] / (B * (T - 3))).float()


def tv_loss(sequence):
    """Total-variation regulariser: mean |x[:,:,1:] - x[:,:,:-1]| for (B, C, T) input."""
    if not sequence.is_cuda:
        raise RuntimeError("tv_loss needs CUDA tensors (no CPU fallback)")
    return _TVFn.apply(sequence)
