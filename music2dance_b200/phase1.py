"""phase1 stick-figure WGAN-GP (BASELINE.json configs[0]) on the libm2d_b200 kernels.

Drop-in for ``phase1/archis/residual.py`` (Generator, Discriminator, LinearBlock: same constructor arguments,
attribute names, state_dict keys and default PyTorch initialisation order) and a fused trainer for the loop
body of ``phase1/train_wgan-gp.py:77-109``: critic iteration (generator forward with train-mode BatchNorm,
gradient penalty ``losses.py:13-54`` in its is_seq=False branch, Wasserstein terms, Adam) and generator update.

Everything is a Linear layer: the contractions run through ``m2d_rowconv`` / ``m2d_wgrad`` (T = 1), BatchNorm,
reductions and Adam through the same kernels as phase3.  The gradient penalty uses the tangent-pass formulation
of wgan.py (the critic is piecewise linear; Dropout is a fixed 0/1 scaling inside one evaluation).  Dropout
masks, generator noise and the interpolation weights are INPUTS (``oracle``-free product code draws them with
torch on the device; the parity tests pass the reference's CPU draws), so results are comparable with the
reference run under the same seed.  LinearBlock's dead ``fc1 -> bn1`` branch (Q1) is reproduced: its BatchNorm
running statistics advance in the generator, its parameters never receive a gradient.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .engine import FlatParams
from .nets import ACT_RELU, BNLayer, Workspace, _conv_from
from .ops import Mat
from .wgan import rows


class LinearBlock(nn.Module):
    """residual.py:51-71 (parameter container; the computation runs in the owning network)."""

    def __init__(self, size, use_bn=False):
        super().__init__()
        self.use_bn, self.size = use_bn, size
        self.fc1 = nn.Linear(size, size, bias=True)
        self.fc2 = nn.Linear(size, size, bias=True)
        if use_bn:
            self.bn1 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
            self.bn2 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        raise RuntimeError("parameter container: use the owning Generator / Discriminator or Phase1Trainer")


class Generator(nn.Module):
    """residual.py:4-24.  forward(noise (B, latent)) -> (B, output_size); inference only (no autograd): training
    goes through Phase1Trainer."""

    def __init__(self, latent_size, size, output_size, nblocks):
        super().__init__()
        self.latent_size, self.size, self.output_size, self.nblocks = latent_size, size, output_size, nblocks
        self.fc1 = nn.Linear(latent_size, size)
        self.bn1 = nn.BatchNorm1d(size, eps=1e-5, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.blocks = nn.Sequential(*[LinearBlock(size, use_bn=True) for _ in range(nblocks)])
        self.dropout = nn.Dropout(p=0.5)
        self.lastfc = nn.Linear(size, output_size)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.phase1.Generator needs CUDA inputs (no CPU fallback)")
        net = _net_of(self, _GenNet)
        B = x.shape[0]
        mask = torch.empty(B, self.size, device=x.device).bernoulli_(0.5) if self.training else None
        with torch.cuda.device(x.device):
            return net.forward(x.detach().float().contiguous(), mask, B, self.training).t.view(B, -1).clone()


class Discriminator(nn.Module):
    """residual.py:27-45.  forward(x (B, 23, 3) | (B, 69)) -> (B, 1); inference only (no autograd)."""

    def __init__(self, input_size, size, nblocks):
        super().__init__()
        self.input_size, self.size, self.nblocks = input_size, size, nblocks
        self.fc1 = nn.Linear(input_size, size)
        self.relu = nn.ReLU(inplace=True)
        self.blocks = nn.Sequential(*[LinearBlock(size) for _ in range(nblocks)])
        self.dropout = nn.Dropout(p=0.5)
        self.lastfc = nn.Linear(size, 1)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.phase1.Discriminator needs CUDA inputs (no CPU fallback)")
        net = _net_of(self, _CriticNet)
        B = x.shape[0]
        mask = (torch.empty(B, self.size, device=x.device).bernoulli_(0.5) if self.training
                else torch.ones(B, self.size, device=x.device) * 0.5 ** 0.5)      # eval: x * m * m * 2 = x
        with torch.cuda.device(x.device):
            X = Mat.of(x.detach().float().reshape(B, -1).contiguous(), 1, B, self.input_size)
            return net.forward(X, Mat.of(mask, 1, B, self.size), B, "inf")["out"].t.view(B, 1).clone()


def _net_of(module, cls):
    net = module.__dict__.get("_m2d_net")
    if net is None or not net.fp.intact():
        net = cls(module)
        module.__dict__["_m2d_net"] = net
    net.ensure_packed()
    return net


class _Net:
    def __init__(self, module):
        ops.check_device(torch.cuda.current_device())
        self.fp = FlatParams(module)
        self.dev = self.fp.device
        self.wk = Workspace(self.dev, scratch_floats=1 << 22)
        self.packed_version = None

    def lin(self, name, Cin, Cout):
        return _conv_from(self.fp.P, self.fp.G, name, Cin, Cout, 1, 1, 0, 1)

    def pack(self):
        for c in self.convs():
            c.pack()
        self.packed_version = self.fp.version()

    def ensure_packed(self):
        if self.fp.version() != self.packed_version:
            self.pack()


class _GenNet(_Net):
    def __init__(self, module):
        super().__init__(module)
        L, S, Oo = module.latent_size, module.size, module.output_size
        self.L, self.S, self.O = L, S, Oo
        P, G = self.fp.P, self.fp.G
        self.fc1, self.bn1 = self.lin("fc1", L, S), BNLayer("bn1", P, G)
        self.blocks = [(self.lin(f"blocks.{i}.fc1", S, S), BNLayer(f"blocks.{i}.bn1", P, G),
                        self.lin(f"blocks.{i}.fc2", S, S), BNLayer(f"blocks.{i}.bn2", P, G))
                       for i in range(module.nblocks)]
        self.last = self.lin("lastfc", S, Oo)
        self.nbt = [b for n, b in module.named_buffers() if n.endswith("num_batches_tracked")]

    def convs(self):
        return [self.fc1, self.last] + [c for d, _, l, _ in self.blocks for c in (d, l)]

    def forward(self, noise, mask, B, train):
        """residual.py:20-24.  noise tensor [B, L]; mask tensor [B, S] of 0/1 or None (eval)."""
        wk, S = self.wk, self.S
        wk.acc_reset()
        z = Mat.of(noise, 1, B, self.L)
        c = wk.mat("g:c0", 1, B, S)
        self.fc1.fwd(z, c, ws=wk.scratch)
        d = wk.mat("g:d0", 1, B, S)
        self.bn1.fwd(c, d, ACT_RELU, train, wk)
        self.sv = [(z, c, d)]
        for i, (dead, bnd, live, bnl) in enumerate(self.blocks):
            cd = wk.mat(f"g:dead{i}", 1, B, S)
            dead.fwd(d, cd, ws=wk.scratch)
            bnd.fwd(cd, None, ACT_RELU, train, wk)                     # Q1: running statistics only
            cl = wk.mat(f"g:c{i + 1}", 1, B, S)
            live.fwd(d, cl, ws=wk.scratch)
            r = wk.mat(f"g:r{i + 1}", 1, B, S)
            bnl.fwd(cl, r, ACT_RELU, train, wk)
            dn = wk.mat(f"g:d{i + 1}", 1, B, S)
            ops.axpby(d, r, dn, B * S, 1.0, 1.0)
            self.sv.append((d, cl, r))
            d = dn
        if mask is not None:
            self.mask = Mat.of(mask, 1, B, S)
            hd = wk.mat("g:hd", 1, B, S)
            ops.mul3(d, self.mask, self.mask, hd, alpha=2.0)           # x * m / (1 - p): m is 0/1, so m*m = m
            d = hd
        else:
            self.mask = None
        fake = wk.mat("g:fake", 1, B, self.O)
        self.last.fwd(d, fake, ws=wk.scratch)
        self.d_last, self.B = d, B
        if train:
            for t in self.nbt:
                t.add_(1)
        return fake

    def backward(self, dfake):
        wk, B, S = self.wk, self.B, self.S
        wk.acc_reset()
        self.last.wgrad(dfake, self.d_last, wk.scratch, acc=wk.acc_slot(self.O))
        e = wk.mat("g:e", 1, B, S)
        self.last.dgrad(dfake, e, ws=wk.scratch)
        if self.mask is not None:
            ops.mul3(e, self.mask, self.mask, e, alpha=2.0)
        for i in range(len(self.blocks) - 1, -1, -1):
            _, _, live, bnl = self.blocks[i]
            d, cl, r = self.sv[i + 1]
            dc = wk.mat(f"g:dc{i + 1}", 1, B, S)
            bnl.bwd(e, r, cl, dc, ACT_RELU, wk)
            live.wgrad(dc, d, wk.scratch, acc=wk.acc_slot(S))
            e2 = wk.mat(f"g:e{i}", 1, B, S)
            live.dgrad(dc, e2, ws=wk.scratch, add=e)
            e = e2
        z, c, d0 = self.sv[0]
        dc = wk.mat("g:dc0", 1, B, S)
        self.bn1.bwd(e, d0, c, dc, ACT_RELU, wk)
        self.fc1.wgrad(dc, z, wk.scratch, acc=wk.acc_slot(S))


class _CriticNet(_Net):
    def __init__(self, module):
        super().__init__(module)
        I, S = module.input_size, module.size
        self.I, self.S = I, S
        self.fc1 = self.lin("fc1", I, S)
        self.blocks = [self.lin(f"blocks.{i}.fc2", S, S) for i in range(module.nblocks)]     # fc1 of a block is dead
        self.last = self.lin("lastfc", S, 1)

    def convs(self):
        return [self.fc1, self.last] + self.blocks

    def forward(self, X, M, n, tag):
        """X Mat [1,n,I]; M Mat [1,n,S] dropout mask (0/1).  residual.py:41-45."""
        wk, S = self.wk, self.S
        h = wk.mat(f"{tag}:h0", 1, n, S)
        self.fc1.fwd(X, h, act=ACT_RELU, ws=wk.scratch)
        sv = {"X": X, "h": [h], "r": [], "M": M}
        for i, fc2 in enumerate(self.blocks):
            r = wk.mat(f"{tag}:r{i}", 1, n, S)
            hn = wk.mat(f"{tag}:h{i + 1}", 1, n, S)
            fc2.fwd(h, hn, act=ACT_RELU, ws=wk.scratch, add=h, y2=r)
            sv["r"].append(r)
            sv["h"].append(hn)
            h = hn
        hd = wk.mat(f"{tag}:hd", 1, n, S)
        ops.mul3(h, M, M, hd, alpha=2.0)
        out = wk.mat(f"{tag}:out", 1, n, 1)
        self.last.fwd(hd, out, ws=wk.scratch)
        sv["hd"], sv["out"] = hd, out
        return sv

    def backward(self, sv, dout, n, tag, dX=None):
        """Backward-data from dout [1,n,1]; returns the pre-activation deltas per layer."""
        wk, S = self.wk, self.S
        e = wk.mat(f"{tag}:e", 1, n, S)
        self.last.dgrad(dout, e, ws=wk.scratch)
        ops.mul3(e, sv["M"], sv["M"], e, alpha=2.0)
        dl = {}
        for i in range(len(self.blocks) - 1, -1, -1):
            d = wk.mat(f"{tag}:d{i + 1}", 1, n, S)
            ops.copy2d(e, d)
            ops.act_bwd(d, sv["r"][i], n * S, ACT_RELU)            # delta of block i's fc2: masked by its ReLU
            dl[i + 1] = d
            e2 = wk.mat(f"{tag}:e{i}", 1, n, S)
            self.blocks[i].dgrad(d, e2, ws=wk.scratch, add=e)
            e = e2
        d0 = wk.mat(f"{tag}:d0", 1, n, S)
        ops.copy2d(e, d0)
        ops.act_bwd(d0, sv["h"][0], n * S, ACT_RELU)
        dl[0] = d0
        if dX is not None:
            self.fc1.dgrad(d0, dX, ws=wk.scratch)
        return dl

    def wgrads(self, dl, dout, X, hs, hd, scale, beta, bias):
        """Weight gradients from deltas dl and layer inputs (X, hs[i] input of block i, hd input of lastfc)."""
        wk = self.wk
        A = lambda c: wk.acc_slot(c.Cout)
        kw = dict(scale=scale, beta=beta, bias=bias)
        self.fc1.wgrad(dl[0], X, wk.scratch, acc=A(self.fc1), **kw)
        for i, fc2 in enumerate(self.blocks):
            fc2.wgrad(dl[i + 1], hs[i], wk.scratch, acc=A(fc2), **kw)
        if dout is not None:
            self.last.wgrad(dout, hd, wk.scratch, acc=A(self.last), **kw)

    def tangent(self, sv, V, n, tag):
        """JVP along V [1,n,I] through the ReLU masks and the dropout mask of `sv` (no biases)."""
        wk, S = self.wk, self.S
        t = wk.mat(f"{tag}:t0", 1, n, S)
        self.fc1.fwd(V, t, bias=False, ws=wk.scratch, mask=sv["h"][0], mask_mode=ACT_RELU)
        ts = [t]
        for i, fc2 in enumerate(self.blocks):
            tn = wk.mat(f"{tag}:t{i + 1}", 1, n, S)
            fc2.fwd(t, tn, bias=False, ws=wk.scratch, mask=sv["r"][i], mask_mode=ACT_RELU, add=t)
            ts.append(tn)
            t = tn
        td = wk.mat(f"{tag}:td", 1, n, S)
        ops.mul3(t, sv["M"], sv["M"], td, alpha=2.0)
        return ts, td


def _sub(sv, b0, b1):
    """Rows [b0, b1) of a critic forward state."""
    return {"X": rows(sv["X"], b0, b1), "h": [rows(h, b0, b1) for h in sv["h"]],
            "r": [rows(r, b0, b1) for r in sv["r"]], "M": rows(sv["M"], b0, b1), "hd": rows(sv["hd"], b0, b1),
            "out": rows(sv["out"], b0, b1)}


class Phase1Trainer:
    """Fused phase1 step (train_wgan-gp.py:77-109) on explicit random inputs."""

    def __init__(self, gen, critic, cfg, batch_size):
        self.cfg, self.B = cfg, batch_size
        self.gen, self.critic = gen, critic
        self.Gn, self.Dn = _net_of(gen, _GenNet), _net_of(critic, _CriticNet)
        self.dev = self.Gn.dev
        f = dict(dtype=torch.float32, device=self.dev)
        self.mD, self.vD = torch.zeros_like(self.Dn.fp.flat), torch.zeros_like(self.Dn.fp.flat)
        self.mG, self.vG = torch.zeros_like(self.Gn.fp.flat), torch.zeros_like(self.Gn.fp.flat)
        self.stepD = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.stepG = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.log = torch.zeros(8, **f)
        self.gp = torch.zeros(1, **f)
        self.fake = None

    def _dev(self, t):
        return t.to(self.dev, torch.float32).contiguous()

    def _adam(self, net, m, v, step, lr):
        ops.adam(net.fp.flat, net.fp.grad, m, v, net.fp.flat.numel(), step, float(lr))
        net.pack()

    def critic_iteration(self, real, noise, mask_g, alpha, masks_d, update=True):
        """train_wgan-gp.py:80-94.  real (B,23,3); noise (B,L); mask_g (B,S); alpha (B,1); masks_d = 3 x (B,S) for
        the interpolates, real and fake evaluations of the critic."""
        B, D, G, cfg = self.B, self.Dn, self.Gn, self.cfg
        with torch.cuda.device(self.dev):
            G.ensure_packed()
            D.ensure_packed()
            fake = G.forward(self._dev(noise), self._dev(mask_g), B, True)
            self.fake = fake.t[:B * G.O].view(B, G.O).clone()
            wk = D.wk
            wk.acc_reset()
            I, S = D.I, D.S
            r = self._dev(real).view(B, I)
            X3 = wk.mat("c:X3", 1, 3 * B, I)
            ops.interp(r, fake, self._dev(alpha).view(-1), X3, B, I)
            ops.axpby(r, None, rows(X3, B, 2 * B), B * I, 1.0, 0.0)
            ops.axpby(fake, None, rows(X3, 2 * B, 3 * B), B * I, 1.0, 0.0)
            M3 = Mat.of(torch.cat([self._dev(m) for m in masks_d]).contiguous(), 1, 3 * B, S)
            sv = D.forward(X3, M3, 3 * B, "c")
            sums = wk.acc_slot(4)
            ops.sum_(rows(sv["out"], B, 2 * B), B, sums[0:1])
            ops.sum_(rows(sv["out"], 2 * B, 3 * B), B, sums[1:2])
            # Wasserstein terms: d/dout = -1/B on the real rows, +1/B on the fake rows (weights and biases, overwrite)
            dd = wk.vec("c:dd", 2 * B)
            ops.fill(dd[:B], B, -1.0 / B)
            ops.fill(dd[B:], B, 1.0 / B)
            ddm = Mat(dd, 1, 2 * B, 1)
            sw = _sub(sv, B, 3 * B)
            dl = D.backward(sw, ddm, 2 * B, "c:w")
            D.wgrads(dl, ddm, sw["X"], sw["h"], sw["hd"], 1.0, 0.0, True)
            # gradient penalty on the interpolates: backward-data -> norms -> tangent pass -> weight gradients
            sg = _sub(sv, 0, B)
            ones = wk.vec("c:ones", B)
            ops.fill(ones, B, 1.0)
            g = wk.mat("c:g", 1, B, I)
            dlg = D.backward(sg, Mat(ones, 1, B, 1), B, "c:gp", dX=g)
            ss = wk.acc_slot(B)
            ops.rows_sumsq(g, B, I, ss)
            k0, k1 = wk.vec("c:k0", B), wk.vec("c:k1", B)
            ops.gp_finalize(ss, None, B, self.gp, k0, k1)
            ops.scale_rows(g, k0, g, B, I)
            ts, td = D.tangent(sg, g, B, "c:gp")
            gamma = float(cfg["gamma"])
            D.wgrads(dlg, None, g, ts, None, gamma, 1.0, False)
            ops.colsum(td, D.last.gw, wk.acc_slot(S), scale=gamma, beta=1.0)
            ops.wgan_scalars(sums, self.gp, B, 1, 1, gamma, 0.0, 0, self.log)
            if update:
                self._adam(D, self.mD, self.vD, self.stepD, cfg["lr_critic"])
            lg = self.log.cpu()
            return dict(loss_critic=float(lg[0]), gp=float(lg[1]), w_dist=float(lg[2]))

    def generator_update(self, real, noise, mask_g, masks_d, update=True):
        """train_wgan-gp.py:97-106.  masks_d = 2 x (B,S): real and fake evaluations of the critic."""
        B, D, G, cfg = self.B, self.Dn, self.Gn, self.cfg
        with torch.cuda.device(self.dev):
            G.ensure_packed()
            D.ensure_packed()
            fake = G.forward(self._dev(noise), self._dev(mask_g), B, True)
            wk = D.wk
            wk.acc_reset()
            I, S = D.I, D.S
            X2 = wk.mat("g:X2", 1, 2 * B, I)
            ops.axpby(self._dev(real).view(B, I), None, rows(X2, 0, B), B * I, 1.0, 0.0)
            ops.axpby(fake, None, rows(X2, B, 2 * B), B * I, 1.0, 0.0)
            M2 = Mat.of(torch.cat([self._dev(m) for m in masks_d]).contiguous(), 1, 2 * B, S)
            sv = D.forward(X2, M2, 2 * B, "g")
            sums = wk.acc_slot(4)
            ops.sum_(rows(sv["out"], 0, B), B, sums[0:1])
            ops.sum_(rows(sv["out"], B, 2 * B), B, sums[1:2])
            dd = wk.vec("g:dd", B)
            ops.fill(dd, B, -1.0 / B)                                  # d(err_real - err_fake)/d out_fake
            dfake = wk.mat("g:dfake", 1, B, I)
            D.backward(_sub(sv, B, 2 * B), Mat(dd, 1, B, 1), B, "g:w", dX=dfake)
            ops.wgan_scalars(sums, None, B, 1, 1, 0.0, 0.0, 1, self.log)
            G.backward(dfake)
            if update:
                self._adam(G, self.mG, self.vG, self.stepG, cfg["lr_gen"])
            lg = self.log.cpu()
            return dict(loss_gen=float(lg[0]))

    def critic_grads(self):
        return {n: self.Dn.fp.G[n].detach().cpu().clone() for n in self.Dn.fp.names}

    def generator_grads(self):
        return {n: self.Gn.fp.G[n].detach().cpu().clone() for n in self.Gn.fp.names}
