"""Drop-in for the hot-path helpers of the reference's ``utils.py``:
``slice_audio_batch`` / ``slice_audio_sequence`` (utils.py:329-353) and
``initialize_weights`` (utils.py:267-313).  Same names, arguments and results;
the windowing runs as a CUDA gather kernel (bit-exact — it is pure indexing)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _cuda_device(device):
    if not torch.cuda.is_available():
        raise RuntimeError("music2dance_b200.utils.slice_audio_batch needs a CUDA device (no CPU fallback)")
    d = torch.device(device) if device is not None else torch.device("cpu")
    return d if d.type == "cuda" else torch.device("cuda", torch.cuda.current_device())


def slice_audio_sequence(seq, audio_feat_samples, cutting_stride, pad_samples, device="cpu"):
    """(A,) -> (n_windows, audio_feat_samples); zero-pad pad//2 left, the rest right."""
    return slice_audio_batch(seq, audio_feat_samples, cutting_stride, pad_samples, device)


def slice_audio_batch(batch, audio_feat_samples, cutting_stride, pad_samples, device="cpu"):
    """1-D or 2-D audio -> overlapping windows, as utils.py:344-353.

    The result lives on the CUDA device (the reference builds it on the CPU and the
    caller moves it with ``.to(device)``, phase3/train.py:189-193 — that move becomes
    a no-op).  The fused trainer never materialises the windows at all: the first
    encoder convolution reads the raw audio with the same index arithmetic."""
    dev = batch.device if batch.is_cuda else _cuda_device(device)
    one = batch.dim() == 1
    a = (batch.unsqueeze(0) if one else batch).to(dev, torch.float32).contiguous()
    nseq, A = a.shape
    nwin = (A + pad_samples - audio_feat_samples) // cutting_stride + 1
    out = torch.empty(nseq, nwin, audio_feat_samples, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        ops.slice_audio(a, out, nseq, A, nwin, audio_feat_samples, cutting_stride, pad_samples // 2)
    return out[0] if one else out


def initialize_weights(net, initialisation=None, bias=None):
    """xavier_normal_ (or normal_(mean, std)) over every Conv / Linear weight and GRU
    weight_* in module-traversal order; biases zeroed only if `bias` is given."""
    def _w(t):
        if initialisation is None:
            nn.init.xavier_normal_(t)
        else:
            nn.init.normal_(t, initialisation[0], initialisation[1])
    for m in net.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear, nn.ConvTranspose1d, nn.ConvTranspose2d)):
            _w(m.weight.data if initialisation is not None else m.weight)
            if bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.GRU):
            for names in m._all_weights:
                for n in names:
                    if "weight" in n:
                        _w(m._parameters[n])


def nparams(model):
    return sum(p.numel() for p in model.parameters())
