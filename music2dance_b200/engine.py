"""Parameter flattening and per-module kernel executors shared by the drop-in
modules (archis/default.py, losses.py) and the fused trainer (trainer.py)."""
from __future__ import annotations

import re
from collections import OrderedDict

import torch

from . import ops
from .nets import CriticNet, GeneratorNet

_DEAD = re.compile(r"decoder\.blocks\.\d+\.(fc1|bn1)\.(weight|bias)$")   # Q1: never receive a gradient


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
        live = [(n, p) for n, p in named if not _DEAD.search(n)]
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
        self.flat = torch.zeros(o, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(o, dtype=torch.float32, device=dev)
        self.P, self.G, self.params = OrderedDict(), OrderedDict(), OrderedDict()
        with torch.no_grad():
            for (n, p), off in zip(self.order, offs):
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.P[n] = view
                self.params[n] = p
                if not _DEAD.search(n):
                    self.G[n] = self.grad[off:off + p.numel()].view(p.shape)
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
