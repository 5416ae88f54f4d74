"""Parameter flattening and per-module kernel executors shared by the drop-in
modules (archis/default.py, losses.py) and the fused trainer (trainer.py)."""
from __future__ import annotations

import os
import re
from collections import OrderedDict

import torch

from . import ops
from .nets import CriticNet, GeneratorNet

# allocator of the gradient buffers (n floats -> zeroed flat tensor).  The data-parallel trainer swaps in the
# symmetric-memory allocator of dp.NvlAllReduce while it builds its engines, so that the kernels write their gradients
# straight into buffers every GPU of the node has mapped.
GRAD_ALLOC = [None]

_DEAD = re.compile(r"decoder\.blocks\.\d+\.(fc1|bn1)\.(weight|bias)$")   # Q1: never receive a gradient


class GradDict(OrderedDict):
    """name -> gradient view in the parameter layout; `.packed[name]` -> the tap-major buffer of the same weight."""

    def __init__(self):
        super().__init__()
        self.packed = {}


class FlatParams:
    """Re-points every parameter of `module` into one flat fp32 buffer (live
    parameters first, the dead LinearBlock branch last) with a matching flat gradient
    buffer, so that Adam and the gradient all-reduce are single flat operations while
    ``state_dict()`` / ``optim.Adam(module.parameters())`` keep working unchanged."""

    def __init__(self, module):
        named = list(module.named_parameters())
        assert named, "module has no parameters"
        dev = named[0][1].device
        assert dev.type == "cuda", "music2dance_b200 runs on CUDA devices only (no CPU fallback)"
        # convolution weights whose gradient the weight-gradient GEMM leaves tap-major (nets.ConvLayer: k > 1 and
        # Cin > 1) come last among the live parameters: [0, n_plain_padded) of the gradient buffer is then everything
        # Adam / the all-reduce read in the parameter layout, and `gpk` (same sizes, same order) holds the rest tap-major
        is_packed = lambda p: p.dim() == 3 and p.shape[1] > 1 and p.shape[2] > 1
        live_all = [(n, p) for n, p in named if not _DEAD.search(n)]
        live = [(n, p) for n, p in live_all if not is_packed(p)] + [(n, p) for n, p in live_all if is_packed(p)]
        n_plain = sum(1 for _, p in live_all if not is_packed(p))
        dead = [(n, p) for n, p in named if _DEAD.search(n)]
        self.order = live + dead
        self.n_live = sum(p.numel() for _, p in live)
        total = sum(p.numel() for _, p in self.order)
        # 16-byte aligned slots so that vectorised loads stay legal
        offs, o = [], 0
        for _, p in self.order:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.n_live_padded = offs[len(live)] if dead else o
        self.n_plain_padded = offs[n_plain] if n_plain < len(live) else self.n_live_padded
        self.flat = torch.zeros(o, dtype=torch.float32, device=dev)
        galloc = GRAD_ALLOC[0] or (lambda n: torch.zeros(n, dtype=torch.float32, device=dev))
        self.grad = galloc(o)
        self.gpk = galloc(max(self.n_live_padded - self.n_plain_padded, 4))
        self.P, self.G, self.params = OrderedDict(), GradDict(), OrderedDict()
        with torch.no_grad():
            for (n, p), off in zip(self.order, offs):
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.P[n] = view
                self.params[n] = p
                if not _DEAD.search(n):
                    self.G[n] = self.grad[off:off + p.numel()].view(p.shape)
                    if is_packed(p):
                        q = off - self.n_plain_padded
                        self.G.packed[n] = self.gpk[q:q + p.numel()]
        # BatchNorm step counters: views of one int64 buffer so that a forward bumps them all at once
        nbt = [(n, b) for n, b in module.named_buffers() if n.endswith("num_batches_tracked")]
        if nbt:
            flat_nbt = torch.zeros(len(nbt), dtype=torch.int64, device=dev)
            with torch.no_grad():
                for i, (n, b) in enumerate(nbt):
                    flat_nbt[i] = b
                    b.data = flat_nbt[i]
            self.P["__nbt_flat__"] = flat_nbt
        for n, b in module.named_buffers():
            self.P[n] = b
        self.names = [n for n, _ in named]             # original parameter order
        self.device = dev

    def intact(self):
        """False if someone moved / replaced the parameters behind our back."""
        for n, p in self.params.items():
            if p.data_ptr() != self.P[n].data_ptr():
                return False
        return True

    def version(self):
        return sum(p._version for p in self.params.values())

    def grad_buffers(self):
        """What an optimiser step reads, as flat tensors (the data-parallel all-reduce runs over exactly these):
        the parameter-layout gradients of everything but the tap-major convolution weights, and the tap-major arena."""
        out = [self.grad[:self.n_plain_padded]]
        if self.n_live_padded > self.n_plain_padded:
            out.append(self.gpk[:self.n_live_padded - self.n_plain_padded])
        return out


class AdamPack:
    """Device table of m2d_adam_pack items for (a subset of) a network's live parameters: convolution / linear weights
    that have packed copies become tiles (gradient read tap-major where the weight-gradient GEMM wrote it that way,
    re-layouts written from the same shared-memory tile); everything else is a plain range.  `m`, `v`: flat Adam moment
    buffers laid out like `fp.flat`.  `only` / `exclude`: lists of ConvLayer objects selecting a subset of the weights
    (plain ranges belong to the table built with only=None unless `flat_range` = (lo, hi) float offsets into the
    flat buffers selects the ones this table owns)."""

    # floats of one output row's slice held in shared memory (x 32 rows): 800 -> 100 KiB tiles, 512-thread blocks,
    # 2 per SM; 400 -> 50 KiB tiles, 256-thread blocks, 4 per SM (M2D_AP_TILE)
    TILE = int(os.environ.get("M2D_AP_TILE", "800"))
    ROWS = int(os.environ.get("M2D_AP_ROWS", "32"))       # output rows per tile (multiple of 4, <= 32)
    FLAT_CHUNK = 4096

    def __init__(self, fp, net, m, v, only=None, exclude=None, flat_range=None):
        import numpy as np
        dev = fp.device
        convs = {c.w.data_ptr(): c for c in net.convs() if c.gw is not None}
        sel = None if only is None else {id(c) for c in only}
        exc = set() if exclude is None else {id(c) for c in exclude}
        base = fp.flat.data_ptr()
        out_dt = np.dtype([("dst", "<u8"), ("dst_tiled", "<u8"), ("kind", "<i4"), ("stride", "<i4"),
                           ("reserved", "<i4"), ("pad_", "<i4")])
        dt = np.dtype([("p", "<u8"), ("m", "<u8"), ("v", "<u8"), ("g", "<u8"), ("flat_n", "<i8"),
                       ("Cout", "<i4"), ("Cin", "<i4"), ("k", "<i4"), ("g_packed", "<i4"),
                       ("co0", "<i4"), ("nco", "<i4"), ("ci0", "<i4"), ("nci", "<i4"), ("t0", "<i4"), ("nt", "<i4"),
                       ("n_pack", "<i4"), ("pad_", "<i4"), ("pk", out_dt, (3,))])
        assert dt.itemsize == 8 * 5 + 4 * 12 + 3 * 32
        items, flats = [], []
        self.smem_floats = 0
        self.nparams, self.bytes = 0, 0       # parameters covered / algorithmic bytes per step (28 B Adam + 12 B per packed layout)
        ptr = lambda t: 0 if t is None else t.data_ptr()
        n_live = fp.n_live_padded
        for name, prm in fp.params.items():
            off = (prm.data_ptr() - base) // 4
            if off >= n_live:
                continue                                          # dead LinearBlock branch (Q1): no gradient, no update
            conv = convs.get(prm.data_ptr())
            if conv is None or prm.dim() < 2:
                if (sel is None and flat_range is None) or (flat_range is not None and flat_range[0] <= off < flat_range[1]):
                    flats.append((off, prm.numel()))
                continue
            if (sel is not None and id(conv) not in sel) or id(conv) in exc:
                continue
            Cout, Cin, k = conv.Cout, conv.Cin, conv.k
            packs = conv.pack_entries()
            assert len(packs) <= 3
            g_packed = conv.gwp is not None
            g_ptr = conv.gwp.data_ptr() if g_packed else fp.grad.data_ptr() + 4 * off
            nt_full = min(k, 25)                                  # taps per tile (odd when split: conflict-free transposes)
            if Cin == 1:
                nci_full = 1
            else:
                nci_full = min(Cin, max(32, (self.TILE // nt_full) // 32 * 32))
            for co0 in range(0, Cout, self.ROWS):
                nco = min(self.ROWS, Cout - co0)
                for ci0 in range(0, Cin, nci_full):
                    nci = min(nci_full, Cin - ci0)
                    for t0 in range(0, k, nt_full):
                        nt = min(nt_full, k - t0)
                        it = np.zeros((), dtype=dt)
                        it["p"], it["m"], it["v"] = prm.data_ptr(), m.data_ptr() + 4 * off, v.data_ptr() + 4 * off
                        it["g"], it["flat_n"] = g_ptr, 0
                        it["Cout"], it["Cin"], it["k"], it["g_packed"] = Cout, Cin, k, int(g_packed)
                        it["co0"], it["nco"], it["ci0"], it["nci"], it["t0"], it["nt"] = co0, nco, ci0, nci, t0, nt
                        it["n_pack"] = len(packs)
                        it["pad_"] = int(all(x % 4 == 0 for x in (Cout, Cin, co0, nco, ci0, nci)))
                        for j, e in enumerate(packs):
                            it["pk"][j] = (ptr(e[1]), ptr(e[2]), e[7], e[6], e[8] if len(e) > 8 else 0, 0)
                        items.append(it)
                        self.smem_floats = max(self.smem_floats, nco * ((nci * nt) | 1))
                        self.nparams += nco * nci * nt
                        self.bytes += nco * nci * nt * (28 + sum((4 if e[1] is not None else 0) + (8 if e[2] is not None else 0)
                                                                 for e in packs))
        # plain ranges: adjacent parameters (16-byte aligned slots, zero padding between them) merge into runs
        flats.sort()
        runs = []
        for off, n in flats:
            end = off + (n + 3) // 4 * 4
            if runs and runs[-1][1] == off:
                runs[-1][1] = end
            else:
                runs.append([off, end])
        for lo, hi in runs:
            for c0 in range(lo, hi, self.FLAT_CHUNK):
                n = min(self.FLAT_CHUNK, hi - c0)
                it = np.zeros((), dtype=dt)
                it["p"], it["m"], it["v"] = base + 4 * c0, m.data_ptr() + 4 * c0, v.data_ptr() + 4 * c0
                it["g"], it["flat_n"] = fp.grad.data_ptr() + 4 * c0, n
                items.append(it)
                self.nparams += n
                self.bytes += 28 * n
        self.n = len(items)
        arr = np.array(items, dtype=dt) if items else np.zeros(0, dtype=dt)
        self.table = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(dev) if self.n else None
        self.counters = torch.zeros(2, dtype=torch.int32, device=dev)      # [0] Adam step count, [1] finished blocks
        self.keep = (m, v)

    def step(self, lr, gscale=1.0):
        if self.n:
            ops.adam_pack(self.table, self.n, self.smem_floats, self.counters, float(lr), gscale=gscale,
                          nbytes=self.bytes)


class Engine:
    """Lazily built executor bound to a drop-in module."""

    def __init__(self, module, kind, cfg):
        ops.check_device(torch.cuda.current_device())
        self.fp = FlatParams(module)
        self.kind, self.cfg = kind, cfg
        Net = GeneratorNet if kind == "gen" else CriticNet
        self.net = Net(self.fp.P, self.fp.G, cfg)
        self.packed_version = None
        self.fid = 0            # forward generation counter (stale-backward guard)
        self.slot = 0
        self.slot_gen = {}

    def ensure_packed(self):
        v = self.fp.version()
        if v != self.packed_version:
            self.net.pack()
            self.packed_version = v

    def mark_dirty(self):
        self.packed_version = None

    def next_slot(self, ring=6):
        self.slot = (self.slot + 1) % ring
        self.slot_gen[self.slot] = self.slot_gen.get(self.slot, 0) + 1
        return self.slot, self.slot_gen[self.slot]

    def grads_in_param_order(self, scale=None):
        self.net.unpack_grads()          # tap-major weight gradients -> parameter layout
        out = []
        for n in self.fp.names:
            g = self.fp.G.get(n)
            if g is None:
                out.append(None)
            else:
                out.append(g.clone() if scale is None else g * scale)
        return out


def engine_of(module, kind, cfg_fn):
    eng = module.__dict__.get("_m2d_engine")
    if eng is None or not eng.fp.intact():
        eng = Engine(module, kind, cfg_fn())
        module.__dict__["_m2d_engine"] = eng
    return eng
