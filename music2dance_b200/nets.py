"""Manual forward/backward executors for the phase3 generator and critic.

These classes own no parameters: they are handed the parameter / gradient tensors
of the drop-in modules (``archis/default.py``) and drive the C-ABI kernels on
channels-last activations.  Reference: phase3/archis/default.py (every class),
losses.py:5-60, phase3/train.py:186-237.
"""
from __future__ import annotations

import os

import torch

from . import ops
from .ops import Mat

ACT_ID, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3

# CUDA stream priorities of the fused step (lower = served first when SMs free up; kernels capture their stream's
# priority into the graph).  The dependent chain of a critic iteration — pose branch / fusion on the capture stream,
# audio branch on its side stream — outranks the generator forwards (needed one iteration later), which outrank the
# weight-gradient GEMMs and re-layouts (leaves of the dependency graph: big grids that would otherwise take every SM
# while a 20-microsecond link of the chain waits).  Measured on B200 at batch 7 (profiles/r02_*): the chain does finish
# 30 % earlier (tangent pass done at 1.12 ms instead of 1.60 ms of an iteration), but the weight gradients then pile up
# behind it and the iteration ends at the same time — the step is bound by total SM time, not by the chain: 56.3 vs
# 57.0 train steps/s.  Hence opt-in (M2D_PRIO=1); default: every stream at the default priority.
_PRIO_ON = os.environ.get("M2D_PRIO", "0") != "0"
PRIO_CHAIN, PRIO_LATE, PRIO_GEN, PRIO_LEAF = (-3, -2, -1, 0) if _PRIO_ON else (0, 0, 0, 0)


# M2D_BN_FUSED=1: train-mode BatchNorm as ONE launch (m2d_bn_train: statistics, grid-wide rendezvous, apply) instead of
# two (colstats + bn_apply).  Measured on B200 at batch 7: 99 launches fewer per step (1163 -> 1064) but 56.9 vs 57.5 train
# steps/s — blocks spinning at the rendezvous hold SM slots the concurrent streams could use — hence opt-in.
_NOISE_BWD_SIDE = os.environ.get("M2D_NOISE_BWD_SIDE", "1") != "0"
_FUSION_ONE_LAUNCH = os.environ.get("M2D_FUSION_FUSED", "1") != "0"
_BN_ONE_LAUNCH = os.environ.get("M2D_BN_FUSED", "0") != "0"


def make_stream(device, priority):
    return torch.cuda.Stream(device=device, priority=priority)
ACT_CODE = {"id": ACT_ID, "relu": ACT_RELU, "tanh": ACT_TANH, "leaky": ACT_LEAKY}


class Workspace:
    """Named, shape-keyed activation buffers (stable addresses across steps, which
    is what CUDA-graph capture needs) + a shared split-K scratch + fp64 accumulators."""

    def __init__(self, device, scratch_floats=1 << 25, acc_doubles=1 << 16):
        self.device = device
        self.bufs = {}
        self.scratch_floats = scratch_floats
        self._scratch = {}                     # split-K scratch of the FP32 kernels, one per CUDA stream
        self.acc = torch.zeros(acc_doubles, dtype=torch.float64, device=device)
        self.acc_used = 0

    @property
    def scratch(self):
        """Scratch of the stream the caller is launching on (concurrent branches never share one)."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        t = self._scratch.get(key)
        if t is None:
            t = torch.empty(self.scratch_floats, dtype=torch.float32, device=self.device)
            self._scratch[key] = t
        return t

    def mat(self, name, nb, rows, cols, ld=None):
        ld = cols if ld is None else ld
        key = (name, nb, rows, cols, ld)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(nb * rows * ld, dtype=torch.float32, device=self.device)
            self.bufs[key] = t
        return Mat(t, nb, rows, cols, ld)

    def vec(self, name, n, dtype=torch.float32):
        key = (name, n, dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.zeros(n, dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t

    def acc_reset(self):
        """Zero the fp64 accumulator arena and restart slot allocation."""
        self.acc.zero_()
        self.acc_used = 0

    def acc_slot(self, n):
        assert self.acc_used + n <= self.acc.numel(), "fp64 accumulator arena exhausted"
        s = self.acc[self.acc_used:self.acc_used + n]
        self.acc_used += n
        return s

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values()) + 4 * self.scratch_floats * len(self._scratch)


def conv_out_len(L, k, s, p):
    return (L + 2 * p - k) // s + 1


class _Fork:
    """`with net.fork():` runs the enclosed launches on the side stream, ordered after everything already
    enqueued on the current stream; `net.join()` makes the current stream wait for them."""

    def __init__(self, side):
        self.side, self.ctx = side, None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


class ConvLayer:
    """Conv1d / Linear as row-convolution GEMMs.  `w` (Cout,Cin,k) or (Cout,Cin),
    `b` (Cout,), `gw`/`gb` same-shaped gradient tensors (views of a flat buffer)."""

    def __init__(self, name, w, b, gw, gb, Cin, Cout, k=1, stride=1, pad=0, Lin=1, need_dgrad=True, gwp=None):
        self.name, self.w, self.b, self.gw, self.gb = name, w, b, gw, gb
        self.Cin, self.Cout, self.k, self.s, self.p, self.Lin = Cin, Cout, k, stride, pad, Lin
        self.Lout = conv_out_len(Lin, k, stride, pad)
        self.full = (pad == 0 and self.Lout == 1)          # collapses the row axis: acts like Linear(k*Cin, Cout)
        self.need_dgrad = need_dgrad
        dev = w.device
        n = Cout * Cin * k
        z = lambda m: torch.zeros(m, dtype=torch.float32, device=dev)
        self.wp = None if (Cin == 1 or k == 1) else torch.empty(n, dtype=torch.float32, device=dev)
        # merged backward-data (all stride residues in one launch): needs Lin % stride == 0
        self.merged = bool(need_dgrad and stride > 1 and not self.full and Lin % stride == 0)
        self.wd = torch.empty(n, dtype=torch.float32, device=dev) if (need_dgrad and Cin > 1 and not self.merged) else None
        # 3xTF32 operand split of the packed weights, pre-tiled in the tensor core's shared-memory image
        # (ops.tiled_floats; padding stays zero): one bulk copy per pipeline stage
        tf = ops.tiled_floats
        self.wpt = z(tf(Cout, k, Cin) if Cin > 1 else tf(Cout, 1, k))
        self.res = []                                       # (rho, Trho, offset, tiled offset) per stride residue
        off = poff = 0
        for rho in range(stride):
            Trho = max(0, -(-(k - rho) // stride))
            self.res.append((rho, Trho, off, poff))
            off += Cin * Cout * Trho
            poff += tf(Cin, Trho, Cout) if Trho > 0 else 0
        # weight gradients of real convolutions are produced tap-major [co, t*Cin + ci] (coalesced stores) and
        # turned into the parameter layout by one batched kernel per network (unpack_entries)
        # (`gwp` given: a slice of the network's tap-major gradient arena, engine.FlatParams.gpk)
        self.gwp = (gwp if gwp is not None else z(n)) if (gw is not None and k > 1 and Cin > 1) else None
        assert gwp is None or (self.gwp is gwp and gwp.numel() == n)
        if self.merged:
            self.cmax = (stride - 1 + pad) // stride
            self.Tm = max(-(-(k - (r + pad) % stride) // stride) + self.cmax - (r + pad) // stride for r in range(stride))
            self.wdm = z(stride * Cin * self.Tm * Cout)
            self.wdmt = z(tf(stride * Cin, self.Tm, Cout))
        if self.wd is not None:                             # Linear / full-length form: [k*Cin rows][Cout]
            self.wdt = z(tf(k * Cin, 1, Cout) if (self.full or k == 1) else poff)
        else:
            self.wdt = None

    def wf(self):
        return self.w if self.wp is None else self.wp

    def pack_entries(self):
        """(w, dst, Cout, Cin, k, stride, kind) rows of the batched re-layout table (ops.pack_batch)."""
        e = [(self.w, self.wp, self.wpt, self.Cout, self.Cin, self.k, 1, ops.PACK_FWD)]
        if self.merged:
            e.append((self.w, self.wdm, self.wdmt, self.Cout, self.Cin, self.k, self.s, ops.PACK_BWD_MERGED, self.p))
        if self.wd is not None:
            if self.full or self.k == 1:
                e.append((self.w, self.wd, self.wdt, self.Cout, self.Cin, self.k, 1, ops.PACK_FULL_BWD))
            else:
                e.append((self.w, self.wd, self.wdt, self.Cout, self.Cin, self.k, self.s, ops.PACK_BWD))
        return e

    def unpack_entries(self):
        if self.gwp is None:
            return []
        return [(self.gwp, self.gw, None, self.Cout, self.Cin, self.k, 1, ops.UNPACK_GRAD)]

    def unpack_grad(self):
        """Single-layer form of the batched gradient unpack (tests)."""
        if self.gwp is not None:
            tab = getattr(self, "_utab", None)
            if tab is None:
                tab = self._utab = ops.pack_table(self.unpack_entries(), self.w.device)
            ops.pack_batch(*tab)

    def pack(self):
        tab = getattr(self, "_tab", None)
        if tab is None:
            tab = self._tab = ops.pack_table(self.pack_entries(), self.w.device)
        ops.pack_batch(*tab)

    # y[b,l,:] = act(conv(x)[b,l,:] + bias)
    def fwd(self, x, y, act=0, bias=True, ws=None, win=None, **epi):
        ops.rowconv(x, self.wf(), y, T=self.k, Cc=self.Cin, N=self.Cout, sr=self.s, roff0=-self.p,
                    droff=1, bias=self.b if bias else None, act=act, ws=ws, win=win,
                    w_tiled=self.wpt.data_ptr(), **epi)

    # dx = conv_transpose(dy) with fused epilogue (mask by act'(prev), residual add, ...)
    def dgrad(self, dy, dx, ws=None, mask=None, mask_mode=0, add=None, add_before_mask=False, y2=None):
        if self.full or self.k == 1:
            assert self.wd is not None
            K = self.k * self.Cin
            xd = dy.flat_rows()
            fl = self.k > 1          # full-length conv: (n, L, C) -> (n, L*C); Linear: rows as they are
            f = (lambda m: None if m is None else (m.flatten_cols().flat_rows() if fl else m.flat_rows()))
            yd = f(dx)
            ops.rowconv(xd, self.wd, yd, T=1, Cc=self.Cout, N=K, ws=ws, mask=f(mask), mask_mode=mask_mode,
                        add=f(add), add_before_mask=add_before_mask, y2=f(y2),
                        w_tiled=self.wdt.data_ptr())
            return
        s = self.s
        dense = all(m is None or (m.ld == m.cols and m.bs == m.rows * m.ld) for m in (dx, mask, add, y2))
        if self.merged and dense and dx.rows == self.Lin:
            # one stride-1 row convolution: coarse row m holds the s fine rows 4m..4m+s-1 as s*Cin columns
            def mv(m):
                if m is None:
                    return None
                v = Mat(m.t, m.nb, m.rows // s, s * m.cols, s * m.ld, m.bs)
                v.ptr = m.ptr
                return v
            ops.rowconv(dy, self.wdm, mv(dx), T=self.Tm, Cc=self.Cout, N=s * self.Cin, sr=1, roff0=self.cmax,
                        droff=-1, ws=ws, mask=mv(mask), mask_mode=mask_mode, add=mv(add),
                        add_before_mask=add_before_mask, y2=mv(y2),
                        w_tiled=self.wdmt.data_ptr())
            return
        assert self.wd is not None, "strided backward-data into a non-dense buffer is not supported for this layer"
        for r0 in range(s):
            rho, c0 = (r0 + self.p) % s, (r0 + self.p) // s
            _, Trho, off, poff = self.res[rho]
            nrows = -(-(dx.rows - r0) // s)
            if nrows <= 0:
                continue
            assert Trho > 0

            def sub(m):
                if m is None:
                    return None
                v = Mat(m.t, m.nb, nrows, m.cols, m.ld * s, m.bs)
                v.ptr = m.ptr + 4 * r0 * m.ld
                return v
            wd = self.wd[off:]
            ops.rowconv(dy, wd, sub(dx), T=Trho, Cc=self.Cout, N=self.Cin, sr=1, roff0=c0, droff=-1,
                        ws=ws, mask=sub(mask), mask_mode=mask_mode, add=sub(add),
                        add_before_mask=add_before_mask, y2=sub(y2),
                        w_tiled=self.wdt.data_ptr() + 4 * poff)

    def wgrad(self, dy, x, ws, scale=1.0, beta=0.0, win=None, bias=True, acc=None, bbeta=None):
        """gw = beta*gw + scale*dW;  gb = bbeta*gb + scale*db (bbeta defaults to beta)."""
        ops.wgrad(dy, x, self.gw if self.gwp is None else self.gwp, Cout=self.Cout, T=self.k, Cc=self.Cin,
                  sr=self.s, roff0=-self.p, droff=1, scale=scale, beta=beta, ws=ws, win=win,
                  packed=self.gwp is not None)
        if bias and self.gb is not None:
            ops.colsum(dy.flat_rows(), self.gb, acc, scale=scale, beta=beta if bbeta is None else bbeta)


class BNLayer:
    def __init__(self, name, P, G):
        self.name = name
        self.gamma, self.beta = P[name + ".weight"], P[name + ".bias"]
        self.rm, self.rv, self.nbt = P[name + ".running_mean"], P[name + ".running_var"], P[name + ".num_batches_tracked"]
        self.ggamma, self.gbeta = G.get(name + ".weight"), G.get(name + ".bias")
        self.C = self.gamma.numel()
        self.mr = torch.empty(2 * self.C, dtype=torch.float32, device=self.gamma.device)

    def fwd(self, c, a, act, train, wk):
        """a = act(BN(c)).  `a` None: only update running statistics (dead branch)."""
        cf = c.flat_rows()
        groups = getattr(wk, "groups", 1)
        if train and groups > 1:
            # `groups` forwards of this module in one pass (GeneratorNet.forward groups=): per-block statistics, the
            # running statistics advance block by block; no backward follows, so mean / rstd are not kept
            acc = wk.acc_slot(2 * self.C * groups)
            ops.colstats(cf, acc, groups=groups)
            ops.bn_apply(cf, None if a is None else a.flat_rows(), acc, self.gamma, self.beta, self.rm,
                         self.rv, None, act, groups=groups)
        elif train and _BN_ONE_LAUNCH:
            acc = wk.acc_slot(2 * self.C + 1)       # sums, sums of squares, rendezvous counter (zeroed with the arena)
            ops.bn_train(cf, None if a is None else a.flat_rows(), acc, self.gamma, self.beta, self.rm,
                         self.rv, self.mr, act)
        elif train:
            acc = wk.acc_slot(2 * self.C)
            ops.colstats(cf, acc)
            ops.bn_apply(cf, None if a is None else a.flat_rows(), acc, self.gamma, self.beta, self.rm,
                         self.rv, self.mr, act)
        elif a is not None:
            ops.bn_eval(cf, a.flat_rows(), self.gamma, self.beta, self.rm, self.rv, act)

    def bwd(self, e_a, a, c, dc, act_mode, wk):
        """dc = d(loss)/d(c) from e_a = d(loss)/d(a)."""
        acc = wk.acc_slot(2 * self.C)
        ops.bn_bwd_reduce(e_a.flat_rows(), a.flat_rows(), c.flat_rows(), self.mr, act_mode, acc)
        ops.bn_bwd_apply(e_a.flat_rows(), a.flat_rows(), c.flat_rows(), dc.flat_rows(), self.mr, self.gamma,
                         act_mode, acc, self.ggamma, self.gbeta)


def _conv_from(P, G, name, Cin, Cout, k, s, p, Lin, need_dgrad=True):
    return ConvLayer(name, P[name + ".weight"], P[name + ".bias"], G.get(name + ".weight"),
                     G.get(name + ".bias"), Cin, Cout, k, s, p, Lin, need_dgrad,
                     gwp=getattr(G, "packed", {}).get(name + ".weight"))


class ConvBNAct:
    """conv -> BatchNorm(train/eval) -> activation, with its backward."""

    def __init__(self, P, G, conv_name, bn_name, Cin, Cout, k, s, p, Lin, act, need_dgrad=True):
        self.conv = _conv_from(P, G, conv_name, Cin, Cout, k, s, p, Lin, need_dgrad)
        self.bn = BNLayer(bn_name, P, G)
        self.act = act
        self.tag = conv_name

    def fwd(self, x, nb, wk, train, win=None, out=None):
        cv = self.conv
        c = wk.mat(self.tag + ":c", nb, cv.Lout, cv.Cout)
        a = out if out is not None else wk.mat(self.tag + ":a", nb, cv.Lout, cv.Cout)
        cv.fwd(x, c, ws=wk.scratch, win=win)
        if out is not None and (a.ld != a.cols):
            # BN kernels take a row stride, so writing into a concat buffer is direct
            pass
        self.bn.fwd(c, a, self.act, train, wk)
        self.x, self.c, self.a = x, c, a
        return a

    def bwd(self, e_a, wk, win=None, e_x=None, leaf=lambda fn: fn(), **dg):
        """e_a: gradient w.r.t. the post-activation output.  Writes parameter grads;
        if e_x is given, the gradient w.r.t. the input (fused epilogue options in dg).
        leaf: runs a launch sequence that nothing in the backward chain depends on (weight / bias gradients) —
        GeneratorNet passes a side-stream runner, the default runs it inline."""
        cv = self.conv
        dc = wk.mat(self.tag + ":dc", self.c.nb, cv.Lout, cv.Cout)
        self.bn.bwd(e_a, self.a, self.c, dc, self.act, wk)
        acc = wk.acc_slot(cv.Cout)
        leaf(lambda: cv.wgrad(dc, self.x, wk.scratch, win=win, acc=acc))      # nothing downstream reads it
        if e_x is not None:
            cv.dgrad(dc, e_x, ws=wk.scratch, **dg)


def _unpack_net_grads(net):
    """Tap-major weight gradients -> parameter layout, ONE launch per network (call once all weight
    gradients of a backward phase have been accumulated, before Adam / all-reduce / autograd hand-off)."""
    tab = getattr(net, "_unpack_tab", None)
    if tab is None:
        entries = [e for c in net.convs() for e in c.unpack_entries()]
        tab = ops.pack_table(entries, net.dev) if entries else (None, 0)
        net._unpack_tab = tab
    if tab[1]:
        ops.pack_batch(tab[0], tab[1])


def _pack_net(net):
    """Refresh every packed weight copy of a network with ONE kernel launch (the table of
    (source, destination, shape, kind) rows lives in device memory and is built once)."""
    tab = getattr(net, "_pack_tab", None)
    if tab is None:
        entries = [e for c in net.convs() for e in c.pack_entries()]
        tab = ops.pack_table(entries, net.dev) if entries else (None, 0)
        net._pack_tab = tab
    if tab[1]:
        ops.pack_batch(tab[0], tab[1])


# ===========================================================================
# generator
# ===========================================================================

class _Encoder:
    def layers(self):
        raise NotImplementedError


class DefaultEncoder(_Encoder):
    """default.py:59-82: conv(1->32,k250,s50,p124)+BN+ReLU; 5x[conv(k4,s2,p1)+BN+ReLU];
    conv(1024->out,k2)+activ.  The first conv reads raw audio with fused windowing."""

    def __init__(self, P, G, cfg):
        p = "audio_enc.model."
        W = cfg["audio_feat_samples"]
        self.blocks = [ConvBNAct(P, G, p + "conv_layers.0", p + "activations.0.0", 1, 32, 250, 50, 124, W,
                                 ACT_RELU, need_dgrad=False)]
        L, C = self.blocks[0].conv.Lout, 32
        for i in range(1, 6):
            self.blocks.append(ConvBNAct(P, G, p + f"conv_layers.{i}", p + f"activations.{i}.0", C, 2 * C, 4,
                                         2, 1, L, ACT_RELU))
            L, C = self.blocks[-1].conv.Lout, 2 * C
        self.head = _conv_from(P, G, p + "conv_layers.6", C, cfg["input_vector_size"], 2, 1, 0, L)
        assert self.head.Lout == 1, "window size incompatible with the default encoder (Q11)"
        self.act = ACT_CODE[cfg["activ"]]

    def convs(self):
        return [b.conv for b in self.blocks] + [self.head]

    def fwd(self, audio, nb, win, wk, train, out):
        x = audio
        for i, blk in enumerate(self.blocks):
            x = blk.fwd(x, nb, wk, train, win=win if i == 0 else None)
        self.head.fwd(x, out, act=self.act, ws=wk.scratch)
        self.hx, self.out = x, out

    def bwd(self, e_out, audio, nb, win, wk, leaf=lambda fn: fn()):
        """e_out: gradient w.r.t. the encoder output [nb,1,out] (modified in place)."""
        if self.act != ACT_ID:
            ops.act_bwd(e_out, self.out, nb * self.head.Cout, self.act)
        acc = wk.acc_slot(self.head.Cout)
        leaf(lambda: self.head.wgrad(e_out, self.hx, wk.scratch, acc=acc))
        e = wk.mat("enc:e_head", nb, self.head.Lin, self.head.Cin)
        self.head.dgrad(e_out, e, ws=wk.scratch)
        for i in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[i]
            if i > 0:
                e_x = wk.mat(f"enc:e{i}", nb, blk.conv.Lin, blk.conv.Cin)
                blk.bwd(e, wk, e_x=e_x, leaf=leaf)
                e = e_x
            else:
                blk.bwd(e, wk, win=win, leaf=leaf)


class WaveGANEncoder(_Encoder):
    """default.py:114-143: 4x[conv(k25,s4,no pad)+BN+ReLU], conv(256->out,k5)+activ."""

    def __init__(self, P, G, cfg):
        p = "audio_enc.model."
        L, C = cfg["audio_feat_samples"], 1
        self.blocks = []
        for i in range(1, 5):
            Co = 32 * (2 ** (i - 1))
            self.blocks.append(ConvBNAct(P, G, p + f"l{i}", p + f"bn{i}", C, Co, 25, 4, 0, L, ACT_RELU,
                                         need_dgrad=(i > 1)))
            L, C = self.blocks[-1].conv.Lout, Co
        self.head = _conv_from(P, G, p + "l5", C, cfg["input_vector_size"], 5, 1, 0, L)
        assert self.head.Lout == 1, "window size incompatible with the WaveGAN encoder"
        self.act = ACT_CODE[cfg["activ"]]

    convs = DefaultEncoder.convs
    fwd = DefaultEncoder.fwd
    bwd = DefaultEncoder.bwd


class UNetEncoder(_Encoder):
    """default.py:85-111,213-246."""

    def __init__(self, P, G, cfg):
        p = "audio_enc.model."
        W = cfg["audio_feat_samples"]
        self.stem = [ConvBNAct(P, G, p + "conv_layers.0", p + "activations.0.0", 1, 32, 160, 4, 79, W,
                               ACT_LEAKY, need_dgrad=False)]
        L = self.stem[0].conv.Lout
        self.stem.append(ConvBNAct(P, G, p + "conv_layers.1", p + "activations.1.0", 32, 64, 4, 2, 1, L, ACT_LEAKY))
        L = self.stem[1].conv.Lout
        self.stem.append(ConvBNAct(P, G, p + "conv_layers.2", p + "activations.2.0", 64, 128, 4, 2, 1, L, ACT_LEAKY))
        L = self.stem[2].conv.Lout
        self.L0 = L
        C = 128
        q = p + "ublock.convblock"
        Ls = [L, L // 2, L // 4, L // 8]
        self.Ls = Ls
        mk = lambda i, cin, Li: ConvBNAct(P, G, q + f"{i}.conv", q + f"{i}.bn", cin, C, 3, 1, 1, Li, ACT_LEAKY)
        self.cb = [None, mk(1, C, Ls[0]), mk(2, C, Ls[1]), mk(3, C, Ls[2]), mk(4, C, Ls[3]),
                   mk(5, 2 * C, Ls[2]), mk(6, 2 * C, Ls[1]), mk(7, 2 * C, Ls[0])]
        assert 2 * Ls[3] == Ls[2] and 2 * Ls[2] == Ls[1] and 2 * Ls[1] == Ls[0], "U-block needs L divisible by 8"
        self.head = _conv_from(P, G, p + "fc", C, cfg["input_vector_size"], L, 1, 0, L)
        assert self.head.Lout == 1
        self.act = ACT_CODE[cfg["activ"]]
        self.C = C

    def convs(self):
        return [b.conv for b in self.stem] + [b.conv for b in self.cb[1:]] + [self.head]

    def fwd(self, audio, nb, win, wk, train, out):
        C, Ls, cb = self.C, self.Ls, self.cb
        x = audio
        for i, blk in enumerate(self.stem):
            x = blk.fwd(x, nb, wk, train, win=win if i == 0 else None)
        # concat buffers: [upsampled | skip], skip written directly by the producing block
        cat5 = wk.mat("unet:cat5", nb, Ls[2], 2 * C)
        cat6 = wk.mat("unet:cat6", nb, Ls[1], 2 * C)
        cat7 = wk.mat("unet:cat7", nb, Ls[0], 2 * C)
        x1 = cb[1].fwd(x, nb, wk, train, out=cat7.cols_slice(C, 2 * C))
        d1 = wk.mat("unet:d1", nb, Ls[1], C)
        ops.maxpool2(x1, d1, nb, Ls[0], C)
        x2 = cb[2].fwd(d1, nb, wk, train, out=cat6.cols_slice(C, 2 * C))
        d2 = wk.mat("unet:d2", nb, Ls[2], C)
        ops.maxpool2(x2, d2, nb, Ls[1], C)
        x3 = cb[3].fwd(d2, nb, wk, train, out=cat5.cols_slice(C, 2 * C))
        d3 = wk.mat("unet:d3", nb, Ls[3], C)
        ops.maxpool2(x3, d3, nb, Ls[2], C)
        x4 = cb[4].fwd(d3, nb, wk, train)
        ops.upsample2(x4, cat5.cols_slice(0, C), nb, Ls[3], C)
        y3 = cb[5].fwd(cat5, nb, wk, train)
        ops.upsample2(y3, cat6.cols_slice(0, C), nb, Ls[2], C)
        y2 = cb[6].fwd(cat6, nb, wk, train)
        ops.upsample2(y2, cat7.cols_slice(0, C), nb, Ls[1], C)
        y1 = cb[7].fwd(cat7, nb, wk, train)
        self.head.fwd(y1, out, act=self.act, ws=wk.scratch)
        self.sv = (x, x1, x2, x3, x4, y3, y2, y1, d1, d2, d3)
        self.out = out

    def bwd(self, e_out, audio, nb, win, wk, leaf=None):
        C, Ls, cb = self.C, self.Ls, self.cb
        x, x1, x2, x3, x4, y3, y2, y1, d1, d2, d3 = self.sv
        if self.act != ACT_ID:
            ops.act_bwd(e_out, self.out, nb * self.head.Cout, self.act)
        self.head.wgrad(e_out, y1, wk.scratch, acc=wk.acc_slot(self.head.Cout))
        e_y1 = wk.mat("unet:e_y1", nb, Ls[0], C)
        self.head.dgrad(e_out, e_y1, ws=wk.scratch)
        e_cat7 = wk.mat("unet:e_cat7", nb, Ls[0], 2 * C)
        cb[7].bwd(e_y1, wk, e_x=e_cat7)
        e_y2 = wk.mat("unet:e_y2", nb, Ls[1], C)
        ops.upsample2_bwd(e_cat7.cols_slice(0, C), e_y2, nb, Ls[1], C, False)
        e_cat6 = wk.mat("unet:e_cat6", nb, Ls[1], 2 * C)
        cb[6].bwd(e_y2, wk, e_x=e_cat6)
        e_y3 = wk.mat("unet:e_y3", nb, Ls[2], C)
        ops.upsample2_bwd(e_cat6.cols_slice(0, C), e_y3, nb, Ls[2], C, False)
        e_cat5 = wk.mat("unet:e_cat5", nb, Ls[2], 2 * C)
        cb[5].bwd(e_y3, wk, e_x=e_cat5)
        e_x4 = wk.mat("unet:e_x4", nb, Ls[3], C)
        ops.upsample2_bwd(e_cat5.cols_slice(0, C), e_x4, nb, Ls[3], C, False)
        e_d3 = wk.mat("unet:e_d3", nb, Ls[3], C)
        cb[4].bwd(e_x4, wk, e_x=e_d3)
        # x3 receives: skip part of cat5 + maxpool backward
        e_x3 = e_cat5.cols_slice(C, 2 * C)
        ops.maxpool2_bwd(x3, e_d3, e_x3, nb, Ls[2], C, True)
        e_d2 = wk.mat("unet:e_d2", nb, Ls[2], C)
        cb[3].bwd(e_x3, wk, e_x=e_d2)
        e_x2 = e_cat6.cols_slice(C, 2 * C)
        ops.maxpool2_bwd(x2, e_d2, e_x2, nb, Ls[1], C, True)
        e_d1 = wk.mat("unet:e_d1", nb, Ls[1], C)
        cb[2].bwd(e_x2, wk, e_x=e_d1)
        e_x1 = e_cat7.cols_slice(C, 2 * C)
        ops.maxpool2_bwd(x1, e_d1, e_x1, nb, Ls[0], C, True)
        e_x = wk.mat("unet:e_x", nb, Ls[0], C)
        cb[1].bwd(e_x1, wk, e_x=e_x)
        e = e_x
        for i in (2, 1):
            blk = self.stem[i]
            e_in = wk.mat(f"unet:e_s{i}", nb, blk.conv.Lin, blk.conv.Cin)
            blk.bwd(e, wk, e_x=e_in)
            e = e_in
        self.stem[0].bwd(e, wk, win=win)


ENCODERS = {"default": DefaultEncoder, "wavegan": WaveGANEncoder, "unet": UNetEncoder}


class GRUStack:
    """torch.nn.GRU(I, H, n_layers, batch_first) — default.py:349-355."""

    def __init__(self, P, G, name, I, H, n_layers):
        self.name, self.I, self.H, self.n = name, I, H, n_layers
        self.ih, self.hh = [], []
        for l in range(n_layers):
            Il = I if l == 0 else H
            self.ih.append(ConvLayer(f"{name}.ih{l}", P[f"{name}.weight_ih_l{l}"], P[f"{name}.bias_ih_l{l}"],
                                     G.get(f"{name}.weight_ih_l{l}"), G.get(f"{name}.bias_ih_l{l}"),
                                     Il, 3 * H, need_dgrad=True))
            self.hh.append((P[f"{name}.weight_hh_l{l}"], P[f"{name}.bias_hh_l{l}"],
                            G.get(f"{name}.weight_hh_l{l}"), G.get(f"{name}.bias_hh_l{l}")))

    def convs(self):
        return list(self.ih)

    def fwd(self, x, out, B, T, wk, save):
        """x [1, B*T, I] rows (b,t); out: Mat [1, B*T, H] possibly a column slice."""
        H = self.H
        self.xs, self.hs, self.saves = [], [], []
        for l in range(self.n):
            gi = wk.mat(f"{self.name}:gi{l}", 1, B * T, 3 * H)
            self.ih[l].fwd(x, gi, ws=wk.scratch)
            h = out if l == self.n - 1 else wk.mat(f"{self.name}:h{l}", 1, B * T, H)
            sv = wk.mat(f"{self.name}:sv{l}", 1, B * T, 4 * H) if save else None
            w_hh, b_hh, _, _ = self.hh[l]
            ops.gru_forward(gi, w_hh, b_hh, h, h.ld, sv, B, T, H)
            self.xs.append(x)
            self.hs.append(h)
            self.saves.append(sv)
            x = h

    def bwd(self, e_out, B, T, wk, e_x=None, leaf=lambda fn: fn()):
        """e_out: Mat gradient w.r.t. the top layer output; e_x: Mat for d/d(input) or None.
        The chain is BPTT kernel -> backward-data of the input projection -> next layer's BPTT kernel; the four
        parameter gradients of a layer are leaves (`leaf`, see ConvBNAct.bwd)."""
        H = self.H
        e = e_out
        for l in range(self.n - 1, -1, -1):
            dgi = wk.mat(f"{self.name}:dgi{l}", 1, B * T, 3 * H)
            dgh = wk.mat(f"{self.name}:dgh{l}", 1, B * T, 3 * H)
            w_hh, _, gw_hh, gb_hh = self.hh[l]
            h = self.hs[l]
            ops.gru_backward(e, e.ld, h, h.ld, self.saves[l], w_hh, dgi, dgh, B, T, H)
            # dW_hh = sum dgh (x) h_{t-1}: rows shifted by one step inside each sequence
            hB = Mat(h.t, B, T, H, h.ld, T * h.ld)
            hB.ptr = h.ptr
            acc1, acc2 = wk.acc_slot(3 * H), wk.acc_slot(3 * H)

            def grads(l=l, dgi=dgi, dgh=dgh, hB=hB, gw_hh=gw_hh, gb_hh=gb_hh, acc1=acc1, acc2=acc2):
                ops.wgrad(dgh.as_rows(B, T), hB, gw_hh, Cout=3 * H, T=1, Cc=H, sr=1, roff0=-1, droff=1,
                          ws=wk.scratch)
                ops.colsum(dgh, gb_hh, acc1)
                self.ih[l].wgrad(dgi, self.xs[l], wk.scratch, acc=acc2)
            if l > 0:
                e = wk.mat(f"{self.name}:e{l}", 1, B * T, H)
                self.ih[l].dgrad(dgi, e, ws=wk.scratch)
            elif e_x is not None:
                self.ih[l].dgrad(dgi, e_x, ws=wk.scratch)
            leaf(grads)


class GeneratorNet:
    """SequenceGenerator.forward (default.py:25-42) and its backward."""

    def __init__(self, P, G, cfg):
        self.cfg = cfg
        self.dev = next(iter(P.values())).device
        self.enc = ENCODERS[cfg["enc_type"]](P, G, cfg)
        self.I, self.Lat, self.Nz = cfg["input_vector_size"], cfg["latent_vector_size"], cfg["noise_size"]
        self.H = self.Lat - self.Nz
        self.rnn = GRUStack(P, G, "audio_rnn.rnn", self.I, self.H, cfg["n_cells"])
        self.nrnn = GRUStack(P, G, "noise_gen.rnn", self.Nz, self.Nz, 1)
        S, O = cfg["size"], cfg["output_size"]
        self.S, self.O = S, O
        self.fc1 = _conv_from(P, G, "decoder.fc1", self.Lat, S, need_dgrad=True, k=1, s=1, p=0, Lin=1)
        self.bn1 = BNLayer("decoder.bn1", P, G)
        self.blocks = []
        for b in range(cfg["nblocks_gen"]):
            q = f"decoder.blocks.{b}."
            dead = ConvLayer(q + "fc1", P[q + "fc1.weight"], P[q + "fc1.bias"], None, None, S, S, need_dgrad=False)
            live = _conv_from(P, G, q + "fc2", S, S, 1, 1, 0, 1)
            self.blocks.append((dead, BNLayer(q + "bn1", P, G), live, BNLayer(q + "bn2", P, G)))
        self.last = _conv_from(P, G, "decoder.lastfc", S, O, 1, 1, 0, 1)
        self.wk = Workspace(self.dev)
        self.nbt = [v for k, v in P.items() if k.endswith("num_batches_tracked")]
        self.nbt_flat = P.get("__nbt_flat__")
        self.par = os.environ.get("M2D_OVERLAP", "1") != "0"
        self.s_noise = make_stream(self.dev, PRIO_GEN) if self.par else None
        self.s_leaf = make_stream(self.dev, PRIO_LEAF) if self.par else None      # parameter gradients of the backward

    def convs(self):
        c = self.enc.convs() + self.rnn.convs() + self.nrnn.convs() + [self.fc1, self.last]
        for dead, _, live, _ in self.blocks:
            c += [dead, live]
        return c

    def pack(self):
        _pack_net(self)

    def unpack_grads(self):
        _unpack_net_grads(self)

    def window(self, T):
        cfg = self.cfg
        pad = cfg["pad_samples"]
        return (T, cfg["cutting_stride"], pad // 2, None, cfg["audio_feat_samples"])

    def forward(self, audio, noise, B, T, train=True, out=None, slices=None, wk=None, groups=1):
        """audio [B, A] raw (fused windowing) or, if `slices` [B,T,W] is given, explicit
        windows (the drop-in module path).  noise [B,T,Nz].  Returns Mat [1,B*T,O].
        groups > 1 (train mode, forward only): the B sequences are `groups` consecutive batches of B / groups — the
        result, BatchNorm batch statistics and running-statistic updates are those of `groups` successive forwards
        (the generator's weights do not change between the critic iterations of a train step, train.py:187-216), in
        one pass over `wk`, a Workspace of its own; nothing is kept for a backward."""
        assert groups == 1 or (train and wk is not None and wk is not self.wk and B % groups == 0)
        save = groups == 1
        wk, cfg = (self.wk if wk is None else wk), self.cfg
        wk.groups = groups
        wk.acc_reset()
        nb = B * T
        if slices is not None:
            src, win = slices, (1, 0, 0, cfg["audio_feat_samples"], cfg["audio_feat_samples"])
        else:
            w = self.window(T)
            src, win = audio, (T, w[1], w[2], audio.shape[-1], w[4])
        self.src, self.win, self.B, self.T = src, win, B, T
        enc_out = wk.mat("g:enc", nb, 1, self.I)
        z = wk.mat("g:z", 1, nb, self.Lat)
        nz = Mat.of(noise, 1, nb, self.Nz)
        # the noise GRU (120 dependent steps on B CTAs) only meets the audio path at the decoder input: side stream
        side = self.s_noise if self.par else None
        cur = torch.cuda.current_stream(self.dev)
        if side is not None:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self.nrnn.fwd(nz, z.cols_slice(self.H, self.Lat), B, T, wk, save=save)
        ops.mark("g:start")
        self.enc.fwd(src, nb, win, wk, train, enc_out)
        ops.mark("g:enc")
        self.rnn.fwd(enc_out.as_rows(1, nb), z.cols_slice(0, self.H), B, T, wk, save=save)
        ops.mark("g:rnn")
        if side is not None:
            cur.wait_stream(side)
        else:
            self.nrnn.fwd(nz, z.cols_slice(self.H, self.Lat), B, T, wk, save=save)
        c = wk.mat("g:dec_c0", 1, nb, self.S)
        self.fc1.fwd(z, c, ws=wk.scratch)
        d = wk.mat("g:dec_d0", 1, nb, self.S)
        self.bn1.fwd(c, d, ACT_RELU, train, wk)
        self.dec = [(z, c, d)]
        for i, (dead, bnd, live, bnl) in enumerate(self.blocks):
            cd = wk.mat(f"g:dec_dead{i}", 1, nb, self.S)
            dead.fwd(d, cd, ws=wk.scratch)
            bnd.fwd(cd, None, ACT_RELU, train, wk)           # Q1: statistics only
            cl = wk.mat(f"g:dec_c{i + 1}", 1, nb, self.S)
            live.fwd(d, cl, ws=wk.scratch)
            r = wk.mat(f"g:dec_r{i + 1}", 1, nb, self.S)
            bnl.fwd(cl, r, ACT_RELU, train, wk)
            dn = wk.mat(f"g:dec_d{i + 1}", 1, nb, self.S)
            ops.axpby(d, r, dn, nb * self.S, 1.0, 1.0)
            self.dec.append((d, cl, r))
            d = dn
        fake = out if out is not None else wk.mat("g:fake", 1, nb, self.O)
        self.last.fwd(d, fake, ws=wk.scratch)
        ops.mark("g:fwd_end")
        self.d_last = d
        if train:
            if self.nbt_flat is not None:
                self.nbt_flat.add_(groups)       # every num_batches_tracked is a view of this buffer
            else:
                for t in self.nbt:
                    t.add_(groups)
        return fake

    def backward(self, dfake):
        """dfake: Mat [1,B*T,O].  Fills every live parameter gradient (overwrite)."""
        wk, B, T = self.wk, self.B, self.T
        nb = B * T
        wk.acc_reset()
        cur = torch.cuda.current_stream(self.dev)
        side = self.s_leaf if self.par else None

        def leaf(fn):
            """Parameter-gradient launches run on a side stream behind the point of the chain that produced their
            operands; the chain (backward-data, BatchNorm backward, BPTT) continues at once."""
            if side is None:
                fn()
                return
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                fn()
        if side is not None:
            side.wait_stream(cur)            # the arena reset above precedes every accumulation on the side stream
        self.last.wgrad(dfake, self.d_last, wk.scratch, acc=wk.acc_slot(self.O))
        e = wk.mat("g:e_d", 1, nb, self.S)
        self.last.dgrad(dfake, e, ws=wk.scratch)
        for i in range(len(self.blocks) - 1, -1, -1):
            dead, bnd, live, bnl = self.blocks[i]
            d, cl, r = self.dec[i + 1]
            dc = wk.mat(f"g:dc{i + 1}", 1, nb, self.S)
            bnl.bwd(e, r, cl, dc, ACT_RELU, wk)
            live.wgrad(dc, d, wk.scratch, acc=wk.acc_slot(self.S))
            e2 = wk.mat(f"g:e_d{i}", 1, nb, self.S)
            live.dgrad(dc, e2, ws=wk.scratch, add=e)
            e = e2
        z, c, d0 = self.dec[0]
        dc = wk.mat("g:dc0", 1, nb, self.S)
        self.bn1.bwd(e, d0, c, dc, ACT_RELU, wk)
        self.fc1.wgrad(dc, z, wk.scratch, acc=wk.acc_slot(self.S))
        e_z = wk.mat("g:e_z", 1, nb, self.Lat)
        self.fc1.dgrad(dc, e_z, ws=wk.scratch)
        ops.mark("gb:dec")
        # the noise GRU's BPTT (120 dependent steps on B CTAs) shares nothing with the audio path below e_z: side
        # stream, like its forward; its parameter gradients follow it there
        noise = self.s_noise if (self.par and _NOISE_BWD_SIDE) else None
        if noise is not None:
            noise.wait_stream(cur)
            with torch.cuda.stream(noise):
                self.nrnn.bwd(e_z.cols_slice(self.H, self.Lat), B, T, wk)
        else:
            self.nrnn.bwd(e_z.cols_slice(self.H, self.Lat), B, T, wk, leaf=leaf)
        e_enc = wk.mat("g:e_enc", 1, nb, self.I)
        self.rnn.bwd(e_z.cols_slice(0, self.H), B, T, wk, e_x=e_enc, leaf=leaf)
        ops.mark("gb:rnn")
        self.enc.bwd(e_enc.as_rows(nb, 1), self.src, nb, self.win, wk, leaf=leaf)
        ops.mark("gb:enc")
        if side is not None:
            cur.wait_stream(side)
        if noise is not None:
            cur.wait_stream(noise)


# ===========================================================================
# critic
# ===========================================================================

class CriticNet:
    """SequenceDiscriminator / AblatedSequenceDiscriminator (default.py:249-346) with the
    passes the WGAN-GP step needs: forward, backward-data, tangent (JVP) and weight grads."""

    def __init__(self, P, G, cfg):
        self.cfg = cfg
        self.dev = next(iter(P.values())).device
        self.ablated = bool(cfg["ablated"])
        O, Ch, code, T = cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"]
        self.O, self.Ch, self.code, self.T = O, Ch, code, T
        k0 = P["stick_d.conv1.weight"].shape[-1]
        self.s_conv1 = _conv_from(P, G, "stick_d.conv1", O, Ch, k0, 1, (k0 - 1) // 2, T)
        self.s_blocks = []
        for b in range(2):
            self.s_blocks.append((_conv_from(P, G, f"stick_d.blocks.{b}.conv1", Ch, Ch, 7, 1, 3, T),
                                  _conv_from(P, G, f"stick_d.blocks.{b}.conv2", Ch, Ch, 7, 1, 3, T)))
        self.s_fconv = _conv_from(P, G, "stick_d.fconv", Ch, code, P["stick_d.fconv.weight"].shape[-1], 1, 0, T)
        assert self.s_fconv.Lout == 1, "critic sequence length is fixed by fconv (Q12)"
        self.act = ACT_CODE[cfg["activ"]]
        self.a_layers = []
        if not self.ablated:
            A = cfg["audio_length"]
            chans = [1, 32, 64, 128, 256, 512]
            L = A
            for i in range(1, 6):
                self.a_layers.append(_conv_from(P, G, f"audio_d.l{i}", chans[i - 1], chans[i], 25, 4, 11, L,
                                                need_dgrad=True))
                L = self.a_layers[-1].Lout
            self.a_l6 = _conv_from(P, G, "audio_d.l6", 512, code, P["audio_d.l6.weight"].shape[-1], 1, 0, L)
            assert self.a_l6.Lout == 1, "critic audio length is fixed by l6 (Q12)"
        self.F = code if self.ablated else 2 * code
        self.fc1 = _conv_from(P, G, "fc1", self.F, 128, 1, 1, 0, 1)
        self.fc2 = _conv_from(P, G, "fc2", 128, 1, 1, 1, 0, 1)
        self.wk = Workspace(self.dev)
        # the audio branch is independent of the pose branch between the inputs and the fusion MLP:
        # it runs on a side stream (fork / join below; CUDA-graph capture turns that into parallel branches)
        self.par = os.environ.get("M2D_OVERLAP", "1") != "0"
        # s_aud: audio branch next to the pose branch; s_w / s_wa: the Wasserstein backward (pose / audio)
        # next to the gradient-penalty passes — in the fused backward: the weight-gradient GEMMs
        self.s_aud, self.s_w, self.s_wa = ((make_stream(self.dev, PRIO_CHAIN), make_stream(self.dev, PRIO_LEAF),
                                            make_stream(self.dev, PRIO_LEAF)) if self.par else (None, None, None))

    def _side(self):
        cur = torch.cuda.current_stream(self.dev)
        return self.s_wa if (self.s_w is not None and cur.cuda_stream == self.s_w.cuda_stream) else self.s_aud

    def fork(self):
        """Audio-branch work of the chain running on the current stream -> its side stream."""
        return _Fork(self._side() if (self.par and not self.ablated) else None)

    def join(self):
        if self.par and not self.ablated:
            torch.cuda.current_stream(self.dev).wait_stream(self._side())

    def fork_w(self):
        """Second chain (Wasserstein backward) next to the chain on the current stream."""
        return _Fork(self.s_w if self.par else None)

    def join_w(self):
        if self.par:
            torch.cuda.current_stream(self.dev).wait_stream(self.s_w)

    def convs(self):
        c = [self.s_conv1] + [x for blk in self.s_blocks for x in blk] + [self.s_fconv]
        if not self.ablated:
            c += self.a_layers + [self.a_l6]
        return c + [self.fc1, self.fc2]

    def pack(self, split=False):
        """Refresh the packed weight copies.  split=True (single-GPU fused trainer): the two big layers that
        the forward reaches last (audio_d.l5 / l6 = 68 % of the critic's parameters) are re-laid out on a side
        stream and audio_fwd waits for them right before l5, so most of the re-layout overlaps the first
        convolutions of the next iteration instead of sitting between Adam and the forward."""
        if not (split and self.can_split_pack()):
            _pack_net(self)
            return
        self.pack_late_fork()
        self.pack_early()

    def can_split_pack(self):
        return self.par and not self.ablated

    def _split_tabs(self):
        if getattr(self, "_pack_tabs", None) is None:
            late = [self.a_layers[4], self.a_l6]
            early = [c for c in self.convs() if c not in late]
            mk = lambda cs: ops.pack_table([e for c in cs for e in c.pack_entries()], self.dev)
            self._pack_tabs = (mk(early), mk(late))
            self.s_pack = make_stream(self.dev, PRIO_LATE)
        return self._pack_tabs

    def pack_early(self):
        """Re-layout of everything but audio_d.l5 / l6 on the current stream."""
        ops.pack_batch(*self._split_tabs()[0])

    def pack_late_fork(self):
        """Re-layout of audio_d.l5 / l6 on the side stream, ordered after the current stream (the optimiser step
        that produced the weights); the next audio_fwd joins it right before l5."""
        tabs = self._split_tabs()
        cur = torch.cuda.current_stream(self.dev)
        self.s_pack.wait_stream(cur)
        with torch.cuda.stream(self.s_pack):
            ops.pack_batch(*tabs[1])
        self._late_pack = True

    def comm_stream(self):
        """Stream of the data-parallel early-bucket all-reduce (next to the last weight-gradient GEMMs)."""
        if getattr(self, "s_comm", None) is None:
            self.s_comm = make_stream(self.dev, PRIO_LATE)
        return self.s_comm

    def late_fork(self, fn):
        """Run `fn` (the optimiser step of audio_d.l5 / l6) on the re-layout side stream, ordered after the current
        stream; the next audio_fwd joins it right before l5 (same protocol as pack_late_fork)."""
        self._split_tabs()
        cur = torch.cuda.current_stream(self.dev)
        self.s_pack.wait_stream(cur)
        with torch.cuda.stream(self.s_pack):
            fn()
        self._late_pack = True

    def unpack_grads(self):
        _unpack_net_grads(self)

    # ---------------------------------------------------------------- pose branch
    def pose_fwd(self, X, n, tag):
        """X Mat [n,T,O] channels-last; sv["code"] = activ(fconv(...)) as a dense [1,n,code]."""
        wk, T, Ch = self.wk, self.T, self.Ch
        code_out = wk.mat(f"{tag}:code_s", 1, n, self.code)
        r0 = wk.mat(f"{tag}:r0", n, T, Ch)
        self.s_conv1.fwd(X, r0, act=ACT_RELU, ws=wk.scratch)
        sv = {"X": X, "r0": r0, "blk": []}
        x = r0
        for b, (c1, c2) in enumerate(self.s_blocks):
            r1 = wk.mat(f"{tag}:b{b}r1", n, T, Ch)
            c1.fwd(x, r1, act=ACT_RELU, ws=wk.scratch)
            r2 = wk.mat(f"{tag}:b{b}r2", n, T, Ch)
            y = wk.mat(f"{tag}:b{b}y", n, T, Ch)
            c2.fwd(r1, y, act=ACT_RELU, ws=wk.scratch, add=x, y2=r2)
            sv["blk"].append((x, r1, r2, y))
            x = y
        x = self._drop(x, sv, n, f"{tag}:yd")
        self.s_fconv.fwd(x, code_out.as_rows(n, 1), act=self.act, ws=wk.scratch)
        ops.mark(f"{tag}:pose_fwd_end")
        sv["y"], sv["code"] = x, code_out
        return sv

    def _drop(self, x, sv, n, name):
        """Hook in front of the full-length convolution: identity here; the phase2 conditional critic applies its
        Dropout mask (phase2/archis/conditional.py:48)."""
        return x

    def _fconv_dgrad(self, sv, d_code, n, dc2, e):
        """e = d/d(last block output), dc2 = e masked by the ReLU of the last block's conv2 — one fused launch."""
        self.s_fconv.dgrad(d_code.as_rows(n, 1), dc2, ws=self.wk.scratch, y2=e, mask=sv["blk"][-1][2],
                           mask_mode=ACT_RELU)

    def pose_bwd(self, sv, d_code, n, tag, scale=1.0, beta=0.0, wgrads=True, dX=None, bbeta=None, pre_act=False):
        """d_code Mat [1,n,code] = gradient w.r.t. the pose code (post-activ; modified in
        place when activ != id; pre_act: already w.r.t. the pre-activation).  Produces
        pre-activation deltas (kept in `sv` for the GP weight gradients), optional weight
        grads (scale/beta) and optional dX."""
        wk, T, Ch = self.wk, self.T, self.Ch
        if self.act != ACT_ID and not pre_act:
            ops.act_bwd(d_code, sv["code"], n * self.code, self.act)
        dl = {"fconv": d_code}
        e = wk.mat(f"{tag}:e_top", n, T, Ch)
        nb = len(self.s_blocks)
        # gradient w.r.t. last block output y; masked copy = delta of its conv2
        dc2 = wk.mat(f"{tag}:b{nb - 1}dc2", n, T, Ch)
        self._fconv_dgrad(sv, d_code, n, dc2, e)
        for b in range(nb - 1, -1, -1):
            x, r1, r2, y = sv["blk"][b]
            c1, c2 = self.s_blocks[b]
            dc1 = wk.mat(f"{tag}:b{b}dc1", n, T, Ch)
            c2.dgrad(dc2, dc1, ws=wk.scratch, mask=r1, mask_mode=ACT_RELU)
            dl[f"b{b}c2"], dl[f"b{b}c1"] = dc2, dc1
            # e_x = e + dgrad_conv1(dc1); masked by the producer of x
            e_x = wk.mat(f"{tag}:e_b{b}", n, T, Ch)
            if b > 0:
                prev_r2 = sv["blk"][b - 1][2]
                dnext = wk.mat(f"{tag}:b{b - 1}dc2", n, T, Ch)
                c1.dgrad(dc1, dnext, ws=wk.scratch, add=e, add_before_mask=True, y2=e_x, mask=prev_r2,
                         mask_mode=ACT_RELU)
                dc2, e = dnext, e_x
            else:
                dc0 = wk.mat(f"{tag}:dc0", n, T, Ch)
                c1.dgrad(dc1, dc0, ws=wk.scratch, add=e, add_before_mask=True, mask=sv["r0"], mask_mode=ACT_RELU)
                dl["conv1"] = dc0
        sv["delta"] = dl
        ops.mark(f"{tag}:pose_bwd_end")
        if dX is not None:
            self.s_conv1.dgrad(dl["conv1"], dX, ws=wk.scratch)
        if wgrads:
            self.pose_wgrads(dl, sv["X"], sv, scale, beta, bias=True, bbeta=bbeta)
        return dl

    def pose_wgrads(self, dl, X, acts, scale, beta, bias, bbeta=None):
        """weight grads from deltas `dl` and layer inputs (X, acts['r0'], blocks, y)."""
        wk = self.wk
        A = lambda c: wk.acc_slot(c.Cout)
        kw = dict(scale=scale, beta=beta, bias=bias, bbeta=bbeta)
        self.s_conv1.wgrad(dl["conv1"], X, wk.scratch, acc=A(self.s_conv1), **kw)
        for b, (c1, c2) in enumerate(self.s_blocks):
            x, r1 = acts["blk"][b][0], acts["blk"][b][1]
            c1.wgrad(dl[f"b{b}c1"], x, wk.scratch, acc=A(c1), **kw)
            c2.wgrad(dl[f"b{b}c2"], r1, wk.scratch, acc=A(c2), **kw)
        n = acts["y"].nb
        self.s_fconv.wgrad(dl["fconv"].as_rows(n, 1), acts["y"], wk.scratch, acc=A(self.s_fconv), **kw)

    def pose_tangent(self, sv, V, n, tag, t_code, inplace=False):
        """JVP of the pose branch along V [n,T,O] with the ReLU masks of `sv` (no biases).
        inplace: the tangent activations overwrite the forward activations they correspond to (t0 -> r0, t1 -> r1,
        ty -> y; the r2 masks stay), so that `sv` afterwards holds the layer inputs of the penalty's weight gradients."""
        wk, T, Ch = self.wk, self.T, self.Ch
        t0 = sv["r0"] if inplace else wk.mat(f"{tag}:t0", n, T, Ch)
        self.s_conv1.fwd(V, t0, bias=False, ws=wk.scratch, mask=sv["r0"], mask_mode=ACT_RELU)
        tv = {"X": V, "r0": t0, "blk": []}
        x = t0
        for b, (c1, c2) in enumerate(self.s_blocks):
            _, r1, r2, y_fw = sv["blk"][b]
            t1 = r1 if inplace else wk.mat(f"{tag}:tb{b}1", n, T, Ch)
            c1.fwd(x, t1, bias=False, ws=wk.scratch, mask=r1, mask_mode=ACT_RELU)
            ty = y_fw if inplace else wk.mat(f"{tag}:tb{b}y", n, T, Ch)
            c2.fwd(t1, ty, bias=False, ws=wk.scratch, mask=r2, mask_mode=ACT_RELU, add=x)
            tv["blk"].append((x, t1, None, ty))
            x = ty
        m = dict(mask=sv["code"].as_rows(n, 1), mask_mode=self.act) if self.act != ACT_ID else {}
        x = self._drop(x, sv, n, f"{tag}:tyd")
        self.s_fconv.fwd(x, t_code.as_rows(n, 1), bias=False, ws=wk.scratch, **m)
        tv["y"] = x
        return tv

    # ---------------------------------------------------------------- audio branch
    def audio_fwd(self, audio, n, tag, dup=False):
        """audio: tensor [n, A] (C=1 channels-last == raw); sv["code"] dense [1,n,code].
        dup: every activation buffer has 2n batch entries and the forward writes BOTH halves (second copy through
        the epilogue's y2 output): the fused backward (wgan.critic_backward_fused) back-propagates the Wasserstein
        and the penalty upstreams as ONE batch of 2n entries through the same ReLU masks, then overwrites the second
        half in place with the tangent activations so that one weight-gradient GEMM per layer covers both terms."""
        wk = self.wk
        code_out = wk.mat(f"{tag}:code_a", 1, n, self.code)
        A = self.cfg["audio_length"]
        x = audio if isinstance(audio, Mat) else Mat.of(audio, n, A, 1)
        sv = {"X": x, "q": [], "q2": []}
        for i, l in enumerate(self.a_layers):
            if i == 4 and getattr(self, "_late_pack", False):
                torch.cuda.current_stream(self.dev).wait_stream(self.s_pack)      # l5 / l6 copies (pack(split=True))
                self._late_pack = False
            if dup:
                q2 = wk.mat(f"{tag}:q2{l.name}", 2 * n, l.Lout, l.Cout)
                q = q2.batch_slice(0, n)
                l.fwd(x, q, act=ACT_RELU, ws=wk.scratch, y2=q2.batch_slice(n, 2 * n))
                sv["q2"].append(q2)
            else:
                q = wk.mat(f"{tag}:q{l.name}", n, l.Lout, l.Cout)
                l.fwd(x, q, act=ACT_RELU, ws=wk.scratch)
            sv["q"].append(q)
            x = q
            ops.mark(f"{tag}:aud_fwd_l{i + 1}")
        self.a_l6.fwd(x, code_out.as_rows(n, 1), act=self.act, ws=wk.scratch)
        ops.mark(f"{tag}:aud_fwd_l6")
        sv["code"] = code_out
        return sv

    def audio_bwd(self, sv, d_code, n, tag, scale=1.0, beta=0.0, wgrads=True, dX=None, bbeta=None, pre_act=False):
        """Backward through the audio branch from d_code [1,n,code] (pre_act: gradient w.r.t. the
        pre-activation of the code).  dX: tensor [n,A] or None."""
        wk = self.wk
        if self.act != ACT_ID and not pre_act:
            ops.act_bwd(d_code, sv["code"], n * self.code, self.act)
        dl = {"l6": d_code}
        q = sv["q"]
        d = wk.mat(f"{tag}:dq5", n, q[4].rows, q[4].cols)
        self.a_l6.dgrad(d_code.as_rows(n, 1), d, ws=wk.scratch, mask=q[4], mask_mode=ACT_RELU)
        ops.mark(f"{tag}:aud_bwd_l6")
        dl[4] = d
        for i in range(4, 0, -1):
            dn = wk.mat(f"{tag}:dq{i}", n, q[i - 1].rows, q[i - 1].cols)
            self.a_layers[i].dgrad(dl[i], dn, ws=wk.scratch, mask=q[i - 1], mask_mode=ACT_RELU)
            ops.mark(f"{tag}:aud_bwd_l{i + 1}")
            dl[i - 1] = dn
        sv["delta"] = dl
        if dX is not None:
            self.l1_dgrad(dl[0], dX, n)
        if wgrads:
            self.audio_wgrads(dl, sv["X"], sv["q"], scale, beta, bias=True, bbeta=bbeta)
        return dl

    def l1_dgrad(self, d0, dX, n):
        """Gradient w.r.t. the raw audio (the penalty's g1): d0 Mat [n, Lout, 32] -> dX tensor or Mat [n, A]."""
        l1, wk = self.a_layers[0], self.wk
        fast = l1.Cout == 32 and l1.k == 25 and l1.s == 4 and l1.Lin % 4 == 0 and l1.p in (0, 11)
        if l1.merged and not fast:                       # tensor-core path: N = stride columns per coarse sample
            l1.dgrad(d0, dX if isinstance(dX, Mat) else Mat(dX, n, l1.Lin, 1), ws=wk.scratch)
        else:                                            # conv_c1.cu: one warp per 4 samples, taps in registers
            ops.conv_dgrad_c1(d0, l1.w, dX, nb=n, Lout=l1.Lout, Cout=l1.Cout, k=l1.k, stride=l1.s,
                              pad=l1.p, Lin=l1.Lin)

    def audio_wgrads(self, dl, X, q, scale, beta, bias, bbeta=None):
        wk = self.wk
        kw = dict(scale=scale, beta=beta, bias=bias, bbeta=bbeta)
        x = X
        for i, l in enumerate(self.a_layers):
            l.wgrad(dl[i], x, wk.scratch, acc=wk.acc_slot(l.Cout), **kw)
            x = q[i]
        n = x.nb
        self.a_l6.wgrad(dl["l6"].as_rows(n, 1), x, wk.scratch, acc=wk.acc_slot(self.a_l6.Cout), **kw)

    def audio_tangent(self, sv, V, n, tag, t_code, inplace=False):
        """inplace: sv["q"][i] are copies of the forward activations that serve as ReLU masks and are overwritten by
        the tangent activations (every output element is read and written by the same thread of the epilogue)."""
        wk = self.wk
        A = self.cfg["audio_length"]
        x = V if isinstance(V, Mat) else Mat.of(V, n, A, 1)
        x0 = x
        tq = []
        for i, l in enumerate(self.a_layers):
            t = sv["q"][i] if inplace else wk.mat(f"{tag}:t{l.name}", n, l.Lout, l.Cout)
            l.fwd(x, t, bias=False, ws=wk.scratch, mask=sv["q"][i], mask_mode=ACT_RELU)
            tq.append(t)
            x = t
        m = dict(mask=sv["code"].as_rows(n, 1), mask_mode=self.act) if self.act != ACT_ID else {}
        self.a_l6.fwd(x, t_code.as_rows(n, 1), bias=False, ws=wk.scratch, **m)
        return {"X": x0, "q": tq}

    # ---------------------------------------------------------------- fusion MLP
    def fusion_fwd(self, sa, n, tag, dd=None):
        """sa Mat [1,n,F] -> scores Mat [1,n,1]; keeps u = relu(fc1(sa)).  One launch (m2d_fusion_mlp); with dd (the
        upstream of the scores, known in advance in the critic step) the same launch also does fusion_bwd and the
        result is returned as (u, d, dh, dsa)."""
        wk = self.wk
        u = wk.mat(f"{tag}:u", 1, n, 128)
        d = wk.mat(f"{tag}:d", 1, n, 1)
        if not _FUSION_ONE_LAUNCH:
            self.fc1.fwd(sa, u, act=ACT_RELU, ws=wk.scratch)
            self.fc2.fwd(u, d, ws=wk.scratch)
            return (u, d) if dd is None else (u, d) + self.fusion_bwd(dd, u, n, tag)
        if dd is None:
            ops.fusion_mlp(sa, self.fc1.w, self.fc1.b, self.fc2.w, self.fc2.b, u, d)
            return u, d
        dh = wk.mat(f"{tag}:dh", 1, n, 128)
        dsa = wk.mat(f"{tag}:dsa", 1, n, self.F)
        ops.fusion_mlp(sa, self.fc1.w, self.fc1.b, self.fc2.w, self.fc2.b, u, d, dd=dd, dh=dh, dx=dsa)
        return u, d, dh, dsa

    def fusion_bwd(self, dd, u, n, tag):
        """dd Mat [1,n,1] -> (dh [1,n,128] masked, dsa [1,n,F])."""
        wk = self.wk
        dh = wk.mat(f"{tag}:dh", 1, n, 128)
        self.fc2.dgrad(dd, dh, ws=wk.scratch, mask=u, mask_mode=ACT_RELU)
        dsa = wk.mat(f"{tag}:dsa", 1, n, self.F)
        self.fc1.dgrad(dh, dsa, ws=wk.scratch)
        return dh, dsa
