"""phase2 CONDITIONAL (dance-type label) sequence WGAN (BASELINE.json configs[2]) on the libm2d_b200
kernels.

Drop-in for ``phase2/archis/conditional.py``: SequenceGenerator (:6-25 — ``nn.Embedding(4, 4)`` label code
concatenated to the noise at every frame, GRU, FrameDecoder with ``Dropout(0.5)`` in front of ``lastfc`` :110-131)
and SequenceDiscriminator (:28-49 — the label code as 4 extra input channels, conv1, TemporalBlocks, ``Dropout(0.5)``,
lastconv) with the reference's constructor arguments, attribute names, state_dict keys and initial weights under the
same seed (the Embedding's normal_ draws come first in both constructors).

``phase2/train_conditional.py`` does not match these modules in the reference (it passes other constructor
arguments and expects an auxiliary-classifier critic returning a tuple, :76-77,130-137), so ``Phase2CondTrainer``
fuses the loop body of the sibling ``phase2/train.py:134-171`` with the labels threaded through both networks and
``gradient_penalty(lambda x: critic(x, labels), ..., is_seq=True, lp=True)`` (losses.py:13-50; lp=True is what
train_conditional.py:127-128 asks for).  Dropout masks, noise and the interpolation weights are INPUTS (the caller
draws them; the parity tests draw them in the reference's RNG order).

Nothing new on the dense side: the label code is written straight into the label columns of the concatenated
operand (``m2d_embed_rows``: lookup + expand + cat in one pass), the row convolutions / GRU input projection simply
see 4 more columns, dropout is one ``m2d_mul3`` either side of the full-length convolution, and the Embedding
gradient is a deterministic segmented reduction of those columns' input gradient (``m2d_embed_grad``).  The penalty
does not reach the Embedding (the critic is piecewise linear in its input, so d‖∂D/∂x‖/d(label code) = 0; autograd
returns exact zeros there too), which the tangent pass reproduces with zero tangents in the label columns.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .nets import ACT_RELU
from .ops import Mat
from .phase2 import FrameDecoder as _FrameDecoder
from .phase2 import NoiseGen, Phase2Trainer, TemporalBlock, _CriticNet, _GenNet, _net_of
from .utils import initialize_weights
from .wgan import rows, slice_pose_saves

N_CLASSES, EMBED = 4, 4


class FrameDecoder(_FrameDecoder):
    """conditional.py:110-131: the phase2 decoder plus ``dropout`` (no parameters) in front of ``lastfc``."""

    def __init__(self, latent_size, size, output_size, nblocks):
        super().__init__(latent_size, size, output_size, nblocks)
        lastfc = self.lastfc
        del self.lastfc                                   # keep the reference's module order: ..., dropout, lastfc
        self.dropout = nn.Dropout(p=0.5)
        self.lastfc = lastfc


def _check_labels(err, what):
    if int(err.item()):
        raise IndexError(f"{what}: label outside [0, {N_CLASSES}) (nn.Embedding index out of range)")


class SequenceGenerator(nn.Module):
    """conditional.py:6-25.  forward(x (B, T, input_size), labels (B,) int64) -> (B*T, output_size); inference only
    (no autograd): training goes through Phase2CondTrainer.  In train mode the decoder's dropout mask is drawn on the
    device (``bernoulli_(0.5)``), like the reference's nn.Dropout."""

    def __init__(self, input_size, latent_size, size, output_size, n_blocks, n_cells=1):
        super().__init__()
        self.input_size, self.latent_size, self.size, self.output_size = input_size, latent_size, size, output_size
        self.n_blocks, self.n_cells = n_blocks, n_cells
        self.embed_label = nn.Embedding(N_CLASSES, EMBED)
        self.noise_gen = NoiseGen(input_size + EMBED, latent_size, n_cells)
        self.decoder = FrameDecoder(latent_size, size, output_size, n_blocks)
        initialize_weights(self)

    def forward(self, x, labels):
        if not (x.is_cuda and labels.is_cuda):
            raise RuntimeError("music2dance_b200.phase2_cond.SequenceGenerator needs CUDA inputs (no CPU fallback)")
        B, T, _ = x.shape
        net = _net_of(self, _CondGenNet)
        with torch.cuda.device(x.device):
            mask = None
            if self.training:
                mask = Mat.of(torch.empty(B * T, self.size, device=x.device).bernoulli_(0.5), 1, B * T, self.size)
            err = torch.zeros(1, dtype=torch.int32, device=x.device)
            xin = net.assemble(x.detach().float().contiguous(), labels.long().contiguous(), B, T, err)
            out = net.forward(xin, B, T, self.training, mask=mask).t[:B * T * self.output_size] \
                .view(B * T, self.output_size).clone()
            _check_labels(err, "SequenceGenerator")
            return out


class SequenceDiscriminator(nn.Module):
    """conditional.py:28-49.  forward(x (B, 69, T), labels (B,)) -> (B, 1); inference only (no autograd)."""

    def __init__(self, channels_in, channels_h, seqlen, init_ker=7, n_blocks=1):
        super().__init__()
        self.channels_in, self.channels_h, self.seqlen, self.n_blocks = channels_in, channels_h, seqlen, n_blocks
        self.embed_label = nn.Embedding(N_CLASSES, EMBED)
        self.conv1 = nn.Conv1d(channels_in + EMBED, channels_h, kernel_size=init_ker, padding=int((init_ker - 1) / 2))
        self.blocks = nn.Sequential(*[TemporalBlock(channels_h, 7) for _ in range(n_blocks)])
        self.lastconv = nn.Conv1d(channels_h, 1, seqlen)
        self.dropout = nn.Dropout(0.5)
        self.relu = nn.ReLU(inplace=True)
        initialize_weights(self)

    def forward(self, x, labels):
        if not (x.is_cuda and labels.is_cuda):
            raise RuntimeError("music2dance_b200.phase2_cond.SequenceDiscriminator needs CUDA inputs (no CPU fallback)")
        B = x.shape[0]
        net = _net_of(self, _CondCriticNet)
        Ci, T, Ch = self.channels_in, net.T, net.Ch
        with torch.cuda.device(x.device):
            wk = net.wk
            Xp = wk.mat("inf:Xp", B, T, Ci)
            ops.transpose_bcl(x.detach().contiguous().float(), Xp, B, Ci, T)
            X = wk.mat("inf:X", B, T, net.O)
            err = torch.zeros(1, dtype=torch.int32, device=x.device)
            net.assemble(Xp, labels.long().contiguous(), X, B, err)
            drop = None
            if self.training:                                           # mask in the reference's (B, C, T) layout
                drop = net.mask_cl(torch.empty(B, Ch, T, device=x.device).bernoulli_(0.5), B, "inf:drop")
            wk.acc_reset()
            net.drop = drop
            out = net.pose_fwd(X, B, "inf")["code"].t[:B].view(B, 1).clone()
            net.drop = None
            _check_labels(err, "SequenceDiscriminator")
            return out


class _CondGenNet(_GenNet):
    EXTRA_IN = EMBED

    def __init__(self, module):
        super().__init__(module)
        self.embed, self.g_embed = self.fp.P["embed_label.weight"], self.fp.G["embed_label.weight"]
        self.In = module.input_size

    def assemble(self, noise, labels, B, T, err=None):
        """cat((x, embed(labels) expanded over T), 2) -> Mat [1, B*T, input + 4]  (conditional.py:19-21)."""
        xin = self.wk.mat("g:xin", 1, B * T, self.I)
        ops.copy2d(Mat.of(noise, 1, B * T, self.In), xin.cols_slice(0, self.In))
        ops.embed_rows(self.embed, labels, xin.cols_slice(self.In, self.I), B, T, N_CLASSES, err)
        return xin

    def backward(self, dfake, labels):
        B, T = self.B, self.T
        e_x = self.wk.mat("g:e_x", 1, B * T, self.I)
        super().backward(dfake, e_x=e_x)
        ops.embed_grad(e_x.cols_slice(self.In, self.I), labels, self.g_embed, B, T, N_CLASSES)


class _CondCriticNet(_CriticNet):
    EXTRA_IN = EMBED

    def __init__(self, module):
        super().__init__(module)
        self.embed, self.g_embed = self.fp.P["embed_label.weight"], self.fp.G["embed_label.weight"]
        self.Ci = module.channels_in
        self.drop = None              # Mat [n, T, Ch] 0/1 mask of the forward / backward / tangent pass in flight

    def assemble(self, poses, labels, X, n, err=None):
        """cat((x, embed(labels) expanded over T), 1), channels-last: X [n, T, 69 + 4]  (conditional.py:45-46)."""
        Xf = X.flat_rows()
        ops.copy2d(poses.flat_rows(), Xf.cols_slice(0, self.Ci))
        ops.embed_rows(self.embed, labels, Xf.cols_slice(self.Ci, self.O), n, self.T, N_CLASSES, err)

    def mask_cl(self, mask_bct, n, name):
        """Dropout mask in the reference's (n, C, T) layout -> channels-last Mat [n, T, C]."""
        m = self.wk.mat(name, n, self.T, self.Ch)
        ops.transpose_bcl(mask_bct.to(self.dev, torch.float32).contiguous(), m, n, self.Ch, self.T)
        return m

    # --- Dropout(0.5) between the last TemporalBlock and lastconv (conditional.py:48): hooks of CriticNet
    def _drop(self, x, sv, n, name):
        if self.drop is None:
            return x
        xd = self.wk.mat(name, n, self.T, self.Ch)
        ops.mul3(x.flat_rows(), self.drop.flat_rows(), self.drop.flat_rows(), xd.flat_rows(), alpha=2.0)
        return xd

    def _fconv_dgrad(self, sv, d_code, n, dc2, e):
        if self.drop is None:
            return super()._fconv_dgrad(sv, d_code, n, dc2, e)
        self.s_fconv.dgrad(d_code.as_rows(n, 1), e, ws=self.wk.scratch)
        ef, m = e.flat_rows(), self.drop.flat_rows()
        ops.mul3(ef, m, m, ef, alpha=2.0)                               # through the dropout
        ops.copy2d(ef, dc2.flat_rows())
        ops.act_bwd(dc2, sv["blk"][-1][2], n * self.T * self.Ch, ACT_RELU)     # delta of the last block's conv2


class Phase2CondTrainer(Phase2Trainer):
    """Fused conditional step: phase2/train.py:134-171 with ``gen(noise, labels)`` / ``critic(x, labels)`` on explicit
    random inputs (noise, dropout masks, alpha).  Adam state, logging and gradient access come from Phase2Trainer."""

    GEN_NET, CRITIC_NET = _CondGenNet, _CondCriticNet

    def _labels(self, labels):
        lab = labels.to(self.dev, torch.int64).contiguous()
        self.err = torch.zeros(1, dtype=torch.int32, device=self.dev)
        return lab

    def _gen_forward(self, noise, labels, mask_g):
        G, B, T = self.Gn, self.B, self.Dn.T
        xin = G.assemble(self._dev(noise), labels, B, T, self.err)
        mask = None if mask_g is None else Mat.of(self._dev(mask_g).view(-1), 1, B * T, G.S)
        return G.forward(xin, B, T, True, mask=mask)

    def critic_iteration(self, real, labels, noise, mask_g, alpha, masks_d, update=True):
        """real (B,T,23,3); labels (B,) int; noise (B,T,input); mask_g (B*T,size); alpha (B,1); masks_d = 3 x
        (B,channels,T) for the interpolates, the real and the fake evaluation of the critic."""
        B, D, cfg = self.B, self.Dn, self.cfg
        T, Ci, Ow = D.T, D.Ci, D.O
        with torch.cuda.device(self.dev):
            self._ensure()
            lab = self._labels(labels)
            fake = self._gen_forward(noise, lab, mask_g)                         # rows (b, t): channels-last poses
            self.fake = fake.t[:B * T * Ci].view(B * T, Ci).clone()
            wk = D.wk
            wk.acc_reset()
            per = T * Ci
            r = self._dev(real).view(B, per)
            P3 = wk.mat("c:P3", 3 * B, T, Ci)
            ops.interp(r, fake, self._dev(alpha).view(-1), P3, B, per)
            ops.axpby(r, None, rows(P3, B, 2 * B), B * per, 1.0, 0.0)
            ops.axpby(fake, None, rows(P3, 2 * B, 3 * B), B * per, 1.0, 0.0)
            X3 = wk.mat("c:X3", 3 * B, T, Ow)
            lab3 = torch.cat([lab, lab, lab])
            D.assemble(P3, lab3, X3, 3 * B, self.err)
            D.drop = drop3 = D.mask_cl(torch.cat([self._dev(m) for m in masks_d]), 3 * B, "c:drop3")
            sv = D.pose_fwd(X3, 3 * B, "c")
            out = sv["code"]                                                     # [1, 3B, 1] critic scores
            sums = wk.acc_slot(4)
            ops.sum_(rows(out, B, 2 * B), B, sums[0:1])
            ops.sum_(rows(out, 2 * B, 3 * B), B, sums[1:2])
            # Wasserstein terms: -1/B on the real rows, +1/B on the fake rows (weights and biases, overwrite)
            dd = wk.vec("c:dd", 2 * B)
            ops.fill(dd[:B], B, -1.0 / B)
            ops.fill(dd[B:], B, 1.0 / B)
            D.drop = rows(drop3, B, 3 * B)
            dXw = wk.mat("c:dXw", 2 * B, T, Ow)
            D.pose_bwd(slice_pose_saves(sv, B, 3 * B), Mat(dd, 1, 2 * B, 1), 2 * B, "c:w", scale=1.0, beta=0.0,
                       wgrads=True, bbeta=0.0, dX=dXw)
            ops.embed_grad(dXw.flat_rows().cols_slice(Ci, Ow), lab3[:2 * B], D.g_embed, 2 * B, T, N_CLASSES)
            # WGAN-LP penalty on the interpolates (w.r.t. the pose columns only: inputs = interpol, losses.py:32-44)
            D.drop = rows(drop3, 0, B)
            ones = wk.vec("c:ones", B)
            ops.fill(ones, B, 1.0)
            sg = slice_pose_saves(sv, 0, B)
            gX = wk.mat("c:gX", B, T, Ow)
            D.pose_bwd(sg, Mat(ones, 1, B, 1), B, "c:gp", wgrads=False, dX=gX)
            g = wk.mat("c:g", B, T, Ci)
            ops.copy2d(gX.flat_rows().cols_slice(0, Ci), g.flat_rows())
            ss = wk.acc_slot(B)
            ops.rows_sumsq(g, B, per, ss)
            k0 = wk.vec("c:k0", B)
            ops.gp_finalize_lp(ss, B, self.gp, k0)
            ops.scale_rows(g, k0, g, B, per)
            V = wk.mat("c:V", B, T, Ow)                                          # tangent: zero in the label columns
            ops.fill(V, B * T * Ow, 0.0)
            ops.copy2d(g.flat_rows(), V.flat_rows().cols_slice(0, Ci))
            t_code = wk.mat("c:t_code", 1, B, 1)
            tv = D.pose_tangent(sg, V, B, "c:gp", t_code)
            gamma = float(cfg["gamma"])
            D.pose_wgrads(sg["delta"], V, tv, gamma, 1.0, bias=False)
            D.drop = None
            ops.wgan_scalars(sums, self.gp, B, 1, 1, gamma, 0.0, 0, self.log)
            D.unpack_grads()
            if update:
                self._adam(D, self.mD, self.vD, self.stepD, cfg["lr_critic"])
            lg = self.log.cpu()
            _check_labels(self.err, "Phase2CondTrainer")
            return dict(loss_critic=float(lg[0]), gp=float(lg[1]), w_dist=float(lg[2]))

    def generator_update(self, real, labels, noise, mask_g, masks_d, update=True):
        """phase2/train.py:159-171 with labels.  masks_d = 2 x (B,channels,T): real and fake evaluation."""
        B, D, G, cfg = self.B, self.Dn, self.Gn, self.cfg
        T, Ci, Ow = D.T, D.Ci, D.O
        with torch.cuda.device(self.dev):
            self._ensure()
            lab = self._labels(labels)
            fake = self._gen_forward(noise, lab, mask_g)
            wk = D.wk
            wk.acc_reset()
            per = T * Ci
            r = self._dev(real).view(B, per)
            P2 = wk.mat("g:P2", 2 * B, T, Ci)
            ops.axpby(r, None, rows(P2, 0, B), B * per, 1.0, 0.0)
            ops.axpby(fake, None, rows(P2, B, 2 * B), B * per, 1.0, 0.0)
            X2 = wk.mat("g:X2", 2 * B, T, Ow)
            D.assemble(P2, torch.cat([lab, lab]), X2, 2 * B, self.err)
            D.drop = drop2 = D.mask_cl(torch.cat([self._dev(m) for m in masks_d]), 2 * B, "g:drop2")
            sv = D.pose_fwd(X2, 2 * B, "g")
            sums = wk.acc_slot(4)
            ops.sum_(rows(sv["code"], 0, B), B, sums[0:1])
            ops.sum_(rows(sv["code"], B, 2 * B), B, sums[1:2])
            dd = wk.vec("g:dd", B)
            ops.fill(dd, B, -1.0 / B)
            D.drop = rows(drop2, B, 2 * B)
            dX = wk.mat("g:dX", B, T, Ow)
            D.pose_bwd(slice_pose_saves(sv, B, 2 * B), Mat(dd, 1, B, 1), B, "g:w", wgrads=False, dX=dX)
            D.drop = None
            dfake = wk.mat("g:dfake", B, T, Ci)
            ops.copy2d(dX.flat_rows().cols_slice(0, Ci), dfake.flat_rows())
            eta = float(cfg["eta"])
            ops.pose_losses(r, fake, dfake, B, T, Ci, 0.0, eta, True, sums[2:4])      # + eta * d tv / d fake
            ops.wgan_scalars(sums, None, B, B * T * Ci, B * (T - 1) * Ci, 0.0, eta, 1, self.log)
            G.backward(dfake.flat_rows(), lab)
            if update:
                self._adam(G, self.mG, self.vG, self.stepG, cfg["lr_gen"])
            lg = self.log.cpu()
            _check_labels(self.err, "Phase2CondTrainer")
            return dict(loss_gen=float(lg[0]), tv=float(lg[2]))
