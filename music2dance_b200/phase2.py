"""phase2 unconditional sequence WGAN-LP (BASELINE.json configs[1]) on the libm2d_b200 kernels.

Drop-in for ``phase2/archis/default.py``: SequenceGenerator (GRU noise generator + FrameDecoder, default.py:5-24,
88-141) and SequenceDiscriminator (conv1 + TemporalBlocks + lastconv, default.py:27-49,144-163) with the reference's
constructor arguments, attribute names, state_dict keys and initial weights under the same seed, plus a fused
trainer for the loop body of ``phase2/train.py:131-171``: critic iteration with the WGAN-LP penalty
(``losses.gradient_penalty(..., is_seq=True, lp=True)``, losses.py:47-50) and generator update with the total-
variation regulariser (eta = 50).

The networks are compositions of what phase3 already runs: the GRU recurrence (gru.cu), the residual FC decoder
(GeneratorNet's decoder incl. LinearBlock's dead branch, Q1) and the critic's pose branch (row convolutions,
backward-data, tangent pass, weight gradients) — here without audio branch and fusion MLP, the code of
``lastconv`` (one channel) IS the critic's score.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .engine import FlatParams
from .nets import ACT_ID, ACT_RELU, BNLayer, ConvLayer, CriticNet, GRUStack, Workspace, _conv_from, _Fork
from .ops import Mat
from .utils import initialize_weights
from .wgan import rows, slice_pose_saves


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the computation runs in music2dance_b200 CUDA kernels")


class NoiseGen(_Holder):
    def __init__(self, input_size, output_size, n_layers):
        super().__init__()
        self.rnn = nn.GRU(input_size, output_size, n_layers, batch_first=True)


class LinearBlock(_Holder):
    def __init__(self, size, use_bn=False):
        super().__init__()
        self.use_bn, self.size = use_bn, size
        self.fc1 = nn.Linear(size, size, bias=True)
        self.fc2 = nn.Linear(size, size, bias=True)
        if use_bn:
            self.bn1 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
            self.bn2 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)


class FrameDecoder(_Holder):
    def __init__(self, latent_size, size, output_size, nblocks):
        super().__init__()
        self.latent_size, self.size, self.output_size, self.nblocks = latent_size, size, output_size, nblocks
        self.fc1 = nn.Linear(latent_size, size)
        self.bn1 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.blocks = nn.Sequential(*[LinearBlock(size, use_bn=True) for _ in range(nblocks)])
        self.lastfc = nn.Linear(size, output_size)


class TemporalBlock(_Holder):
    def __init__(self, channels, ksize):
        super().__init__()
        self.channels, self.ksize, self.pad = channels, ksize, int((ksize - 1) / 2)
        self.conv1 = nn.Conv1d(channels, channels, kernel_size=ksize, padding=self.pad, dilation=1)
        self.conv2 = nn.Conv1d(channels, channels, kernel_size=ksize, padding=self.pad, dilation=1)
        self.relu = nn.ReLU(inplace=True)


class SequenceGenerator(nn.Module):
    """default.py:5-24.  forward(noise (B, T, input_size), lengths) -> (B*T, output_size); inference only (no
    autograd): training goes through Phase2Trainer."""

    def __init__(self, input_size, latent_size, size, output_size, n_blocks, n_cells=1, device="cpu"):
        super().__init__()
        self.input_size, self.latent_size, self.size, self.output_size = input_size, latent_size, size, output_size
        self.n_blocks, self.n_cells = n_blocks, n_cells
        self.noise_gen = NoiseGen(input_size, latent_size, n_cells)
        self.decoder = FrameDecoder(latent_size, size, output_size, n_blocks)
        initialize_weights(self)
        self.to(device)

    def forward(self, x, lengths):
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.phase2.SequenceGenerator needs CUDA inputs (no CPU fallback)")
        B, T, _ = x.shape
        assert all(int(l) == T for l in lengths), "equal lengths only (the reference trains on fixed 120-frame crops)"
        net = _net_of(self, _GenNet)
        with torch.cuda.device(x.device):
            return net.forward(x.detach().float().contiguous(), B, T, self.training).t[:B * T * self.output_size] \
                .view(B * T, self.output_size).clone()


class SequenceDiscriminator(nn.Module):
    """default.py:27-49.  forward(x (B, 69, T)) -> (B, 1); inference only (no autograd)."""

    def __init__(self, channels_in, channels_h, seqlen, init_ker=7, n_blocks=1, device="cpu"):
        super().__init__()
        self.channels_in, self.channels_h, self.seqlen, self.n_blocks = channels_in, channels_h, seqlen, n_blocks
        self.conv1 = nn.Conv1d(channels_in, channels_h, kernel_size=init_ker, padding=int((init_ker - 1) / 2))
        self.blocks = nn.Sequential(*[TemporalBlock(channels_h, 7) for _ in range(n_blocks)])
        self.lastconv = nn.Conv1d(channels_h, 1, seqlen)
        self.relu = nn.ReLU(inplace=True)
        initialize_weights(self)
        self.to(device)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.phase2.SequenceDiscriminator needs CUDA inputs (no CPU fallback)")
        B = x.shape[0]
        net = _net_of(self, _CriticNet)
        with torch.cuda.device(x.device):
            X = net.wk.mat("inf:X", B, net.T, net.O)
            ops.transpose_bcl(x.detach().contiguous().float(), X, B, net.O, net.T)
            net.wk.acc_reset()
            return net.pose_fwd(X, B, "inf")["code"].t[:B].view(B, 1).clone()


def _net_of(module, cls):
    net = module.__dict__.get("_m2d_net")
    if net is None or not net.fp.intact():
        net = cls(module)
        module.__dict__["_m2d_net"] = net
    if net.fp.version() != net.packed_version:
        net.pack()
    return net


class _GenNet:
    """GRU stack + FrameDecoder (the decoder code mirrors nets.GeneratorNet)."""

    EXTRA_IN = 0                  # columns concatenated to the noise in front of the GRU (conditional: label code)

    def __init__(self, module):
        ops.check_device(torch.cuda.current_device())
        self.fp = FlatParams(module)
        P, G = self.fp.P, self.fp.G
        self.dev = self.fp.device
        self.I, self.H, self.S, self.O = module.input_size + self.EXTRA_IN, module.latent_size, module.size, \
            module.output_size
        S = self.S
        self.rnn = GRUStack(P, G, "noise_gen.rnn", self.I, self.H, module.n_cells)
        self.fc1 = _conv_from(P, G, "decoder.fc1", self.H, S, 1, 1, 0, 1)
        self.bn1 = BNLayer("decoder.bn1", P, G)
        self.blocks = []
        for b in range(module.n_blocks):
            q = f"decoder.blocks.{b}."
            dead = ConvLayer(q + "fc1", P[q + "fc1.weight"], P[q + "fc1.bias"], None, None, S, S, need_dgrad=False)
            self.blocks.append((dead, BNLayer(q + "bn1", P, G), _conv_from(P, G, q + "fc2", S, S, 1, 1, 0, 1),
                                BNLayer(q + "bn2", P, G)))
        self.last = _conv_from(P, G, "decoder.lastfc", S, self.O, 1, 1, 0, 1)
        self.wk = Workspace(self.dev, scratch_floats=1 << 23)
        self.nbt_flat = P.get("__nbt_flat__")
        self.packed_version = None

    def convs(self):
        return self.rnn.convs() + [self.fc1, self.last] + [c for d, _, l, _ in self.blocks for c in (d, l)]

    def pack(self):
        for c in self.convs():
            c.pack()
        self.packed_version = self.fp.version()

    def forward(self, noise, B, T, train, mask=None):
        """noise: tensor (B, T, I) or an already assembled Mat [1, B*T, I]; mask: Mat [1, B*T, S] 0/1 dropout mask in
        front of ``lastfc`` (phase2/archis/conditional.py:127; None = no dropout)."""
        wk, S, nb = self.wk, self.S, B * T
        wk.acc_reset()
        z = wk.mat("g:z", 1, nb, self.H)
        self.rnn.fwd(noise if isinstance(noise, Mat) else Mat.of(noise, 1, nb, self.I), z, B, T, wk, save=True)
        c = wk.mat("g:c0", 1, nb, S)
        self.fc1.fwd(z, c, ws=wk.scratch)
        d = wk.mat("g:d0", 1, nb, S)
        self.bn1.fwd(c, d, ACT_RELU, train, wk)
        self.dec = [(z, c, d)]
        for i, (dead, bnd, live, bnl) in enumerate(self.blocks):
            cd = wk.mat(f"g:dead{i}", 1, nb, S)
            dead.fwd(d, cd, ws=wk.scratch)
            bnd.fwd(cd, None, ACT_RELU, train, wk)                     # Q1: running statistics only
            cl = wk.mat(f"g:c{i + 1}", 1, nb, S)
            live.fwd(d, cl, ws=wk.scratch)
            r = wk.mat(f"g:r{i + 1}", 1, nb, S)
            bnl.fwd(cl, r, ACT_RELU, train, wk)
            dn = wk.mat(f"g:d{i + 1}", 1, nb, S)
            ops.axpby(d, r, dn, nb * S, 1.0, 1.0)
            self.dec.append((d, cl, r))
            d = dn
        fake = wk.mat("g:fake", 1, nb, self.O)
        self.mask = mask
        if mask is not None:
            hd = wk.mat("g:hd", 1, nb, S)
            ops.mul3(d, mask, mask, hd, alpha=2.0)                     # x * m / (1 - p): m is 0/1, so m*m = m
            d = hd
        self.last.fwd(d, fake, ws=wk.scratch)
        self.d_last, self.B, self.T = d, B, T
        if train and self.nbt_flat is not None:
            self.nbt_flat.add_(1)
        return fake

    def backward(self, dfake, e_x=None):
        """e_x: optional Mat [1, B*T, I] receiving the gradient w.r.t. the GRU input (label-code columns)."""
        wk, B, T, S = self.wk, self.B, self.T, self.S
        nb = B * T
        wk.acc_reset()
        self.last.wgrad(dfake, self.d_last, wk.scratch, acc=wk.acc_slot(self.O))
        e = wk.mat("g:e", 1, nb, S)
        self.last.dgrad(dfake, e, ws=wk.scratch)
        if self.mask is not None:
            ops.mul3(e, self.mask, self.mask, e, alpha=2.0)
        for i in range(len(self.blocks) - 1, -1, -1):
            _, _, live, bnl = self.blocks[i]
            d, cl, r = self.dec[i + 1]
            dc = wk.mat(f"g:dc{i + 1}", 1, nb, S)
            bnl.bwd(e, r, cl, dc, ACT_RELU, wk)
            live.wgrad(dc, d, wk.scratch, acc=wk.acc_slot(S))
            e2 = wk.mat(f"g:e{i}", 1, nb, S)
            live.dgrad(dc, e2, ws=wk.scratch, add=e)
            e = e2
        z, c, d0 = self.dec[0]
        dc = wk.mat("g:dc0", 1, nb, S)
        self.bn1.bwd(e, d0, c, dc, ACT_RELU, wk)
        self.fc1.wgrad(dc, z, wk.scratch, acc=wk.acc_slot(S))
        e_z = wk.mat("g:e_z", 1, nb, self.H)
        self.fc1.dgrad(dc, e_z, ws=wk.scratch)
        self.rnn.bwd(e_z, B, T, wk, e_x=e_x)
        for cv in self.convs():
            cv.unpack_grad()


class _CriticNet(CriticNet):
    """conv1 + n TemporalBlocks + lastconv = the pose branch of nets.CriticNet with a one-channel code and no
    fusion MLP; inherits pose_fwd / pose_bwd / pose_wgrads / pose_tangent."""

    EXTRA_IN = 0                  # input channels concatenated to the poses (conditional: label code)

    def __init__(self, module):                                    # noqa: super().__init__ intentionally not called
        ops.check_device(torch.cuda.current_device())
        self.fp = FlatParams(module)
        P, G = self.fp.P, self.fp.G
        self.dev = self.fp.device
        self.cfg, self.ablated, self.par = {}, True, False
        Oo, Ch, T = module.channels_in + self.EXTRA_IN, module.channels_h, module.seqlen
        self.O, self.Ch, self.code, self.T = Oo, Ch, 1, T
        k0 = P["conv1.weight"].shape[-1]
        self.s_conv1 = _conv_from(P, G, "conv1", Oo, Ch, k0, 1, (k0 - 1) // 2, T)
        self.s_blocks = [(_conv_from(P, G, f"blocks.{b}.conv1", Ch, Ch, 7, 1, 3, T),
                          _conv_from(P, G, f"blocks.{b}.conv2", Ch, Ch, 7, 1, 3, T)) for b in range(module.n_blocks)]
        self.s_fconv = _conv_from(P, G, "lastconv", Ch, 1, T, 1, 0, T)
        self.act = ACT_ID
        self.a_layers, self.F = [], 1
        self.wk = Workspace(self.dev, scratch_floats=1 << 24)
        self.s_aud = self.s_w = self.s_wa = None
        self.packed_version = None

    def convs(self):
        return [self.s_conv1] + [x for blk in self.s_blocks for x in blk] + [self.s_fconv]

    def pack(self, split=False):
        for c in self.convs():
            c.pack()
        self.packed_version = self.fp.version()

    def fork(self):
        return _Fork(None)

    def join(self):
        pass

    def unpack_grads(self):
        for c in self.convs():
            c.unpack_grad()


class Phase2Trainer:
    """Fused phase2 step (phase2/train.py:131-171) on explicit random inputs (noise, alpha)."""

    GEN_NET, CRITIC_NET = _GenNet, _CriticNet
    # phase2/train.py:88-89,179-180: both Adam optimisers sit behind MultiStepLR(milestones, gamma=0.8), stepped once
    # per generator update (after the optimiser step)
    LR_MILESTONES, LR_GAMMA = (10000, 35000, 50000), 0.8

    def __init__(self, gen, critic, cfg, batch_size):
        self.cfg, self.B = cfg, batch_size
        self.gen, self.critic = gen, critic
        self.Gn, self.Dn = _net_of(gen, self.GEN_NET), _net_of(critic, self.CRITIC_NET)
        self.dev = self.Gn.dev
        f = dict(dtype=torch.float32, device=self.dev)
        nD, nG = self.Dn.fp.n_live_padded, self.Gn.fp.n_live_padded
        self.mD, self.vD = torch.zeros(nD, **f), torch.zeros(nD, **f)
        self.mG, self.vG = torch.zeros(nG, **f), torch.zeros(nG, **f)
        self.stepD = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.stepG = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.log, self.gp = torch.zeros(8, **f), torch.zeros(1, **f)
        self.fake = None
        self.sched_steps = 0                       # scheduler.step() calls so far (= generator updates applied)

    def _dev(self, t):
        return t.to(self.dev, torch.float32).contiguous()

    def lr_factor(self):
        """MultiStepLR: gamma ** (number of milestones <= scheduler steps taken)."""
        return self.LR_GAMMA ** sum(1 for ms in self.LR_MILESTONES if ms <= self.sched_steps)

    def _adam(self, net, m, v, step, lr):
        n = net.fp.n_live_padded
        ops.adam(net.fp.flat, net.fp.grad, m, v, n, step, float(lr) * self.lr_factor())
        net.pack()
        if net is self.Gn:
            self.sched_steps += 1                  # scheduler_critic.step(); scheduler_gen.step() (train.py:179-180)

    def _ensure(self):
        for net in (self.Gn, self.Dn):
            if net.fp.version() != net.packed_version:
                net.pack()

    def critic_iteration(self, real, noise, alpha, update=True):
        """phase2/train.py:134-154.  real (B,T,23,3); noise (B,T,input); alpha (B,1)."""
        B, D, G, cfg = self.B, self.Dn, self.Gn, self.cfg
        T, Oo = D.T, D.O
        with torch.cuda.device(self.dev):
            self._ensure()
            fake = G.forward(self._dev(noise), B, T, True)                      # rows (b, t): channels-last poses
            self.fake = fake.t[:B * T * Oo].view(B * T, Oo).clone()
            wk = D.wk
            wk.acc_reset()
            per = T * Oo
            r = self._dev(real).view(B, per)
            X3 = wk.mat("c:X3", 3 * B, T, Oo)
            ops.interp(r, fake, self._dev(alpha).view(-1), X3, B, per)
            ops.axpby(r, None, rows(X3, B, 2 * B), B * per, 1.0, 0.0)
            ops.axpby(fake, None, rows(X3, 2 * B, 3 * B), B * per, 1.0, 0.0)
            sv = D.pose_fwd(X3, 3 * B, "c")
            out = sv["code"]                                                     # [1, 3B, 1] critic scores
            sums = wk.acc_slot(4)
            ops.sum_(rows(out, B, 2 * B), B, sums[0:1])
            ops.sum_(rows(out, 2 * B, 3 * B), B, sums[1:2])
            # Wasserstein terms: -1/B on the real rows, +1/B on the fake rows (weights and biases, overwrite)
            dd = wk.vec("c:dd", 2 * B)
            ops.fill(dd[:B], B, -1.0 / B)
            ops.fill(dd[B:], B, 1.0 / B)
            D.pose_bwd(slice_pose_saves(sv, B, 3 * B), Mat(dd, 1, 2 * B, 1), 2 * B, "c:w", scale=1.0, beta=0.0,
                       wgrads=True, bbeta=0.0)
            # WGAN-LP penalty on the interpolates
            ones = wk.vec("c:ones", B)
            ops.fill(ones, B, 1.0)
            sg = slice_pose_saves(sv, 0, B)
            g = wk.mat("c:g", B, T, Oo)
            D.pose_bwd(sg, Mat(ones, 1, B, 1), B, "c:gp", wgrads=False, dX=g)
            ss = wk.acc_slot(B)
            ops.rows_sumsq(g, B, per, ss)
            k0 = wk.vec("c:k0", B)
            ops.gp_finalize_lp(ss, B, self.gp, k0)
            ops.scale_rows(g, k0, g, B, per)
            t_code = wk.mat("c:t_code", 1, B, 1)
            tv = D.pose_tangent(sg, g, B, "c:gp", t_code)
            gamma = float(cfg["gamma"])
            D.pose_wgrads(sg["delta"], g, tv, gamma, 1.0, bias=False)
            ops.wgan_scalars(sums, self.gp, B, 1, 1, gamma, 0.0, 0, self.log)
            D.unpack_grads()
            if update:
                self._adam(D, self.mD, self.vD, self.stepD, cfg["lr_critic"])
            lg = self.log.cpu()
            return dict(loss_critic=float(lg[0]), gp=float(lg[1]), w_dist=float(lg[2]))

    def generator_update(self, real, noise, update=True):
        """phase2/train.py:159-171."""
        B, D, G, cfg = self.B, self.Dn, self.Gn, self.cfg
        T, Oo = D.T, D.O
        with torch.cuda.device(self.dev):
            self._ensure()
            fake = G.forward(self._dev(noise), B, T, True)
            wk = D.wk
            wk.acc_reset()
            per = T * Oo
            r = self._dev(real).view(B, per)
            X2 = wk.mat("g:X2", 2 * B, T, Oo)
            ops.axpby(r, None, rows(X2, 0, B), B * per, 1.0, 0.0)
            ops.axpby(fake, None, rows(X2, B, 2 * B), B * per, 1.0, 0.0)
            sv = D.pose_fwd(X2, 2 * B, "g")
            sums = wk.acc_slot(4)
            ops.sum_(rows(sv["code"], 0, B), B, sums[0:1])
            ops.sum_(rows(sv["code"], B, 2 * B), B, sums[1:2])
            dd = wk.vec("g:dd", B)
            ops.fill(dd, B, -1.0 / B)
            dfake = wk.mat("g:dfake", B, T, Oo)
            D.pose_bwd(slice_pose_saves(sv, B, 2 * B), Mat(dd, 1, B, 1), B, "g:w", wgrads=False, dX=dfake)
            eta = float(cfg["eta"])
            ops.pose_losses(r, fake, dfake, B, T, Oo, 0.0, eta, True, sums[2:4])      # + eta * d tv / d fake
            ops.wgan_scalars(sums, None, B, B * T * Oo, B * (T - 1) * Oo, 0.0, eta, 1, self.log)
            G.backward(dfake.flat_rows())
            if update:
                self._adam(G, self.mG, self.vG, self.stepG, cfg["lr_gen"])
            lg = self.log.cpu()
            return dict(loss_gen=float(lg[0]), tv=float(lg[2]))

    def _grads(self, net):
        out = {}
        for n in net.fp.names:
            g = net.fp.G.get(n)
            out[n] = torch.zeros_like(net.fp.P[n]).cpu() if g is None else g.detach().cpu().clone()
        return out

    def critic_grads(self):
        return self._grads(self.Dn)

    def generator_grads(self):
        return self._grads(self.Gn)
