"""Thin tensor-level wrappers over the C ABI (include/m2d.h).

PyTorch is plumbing here: it owns device memory and the current CUDA stream.
Every function enqueues hand-written kernels from libm2d_b200.so on
``torch.cuda.current_stream()``; none of them falls back to a torch op.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import ACT, RowConvArgs, WgradArgs, call

LAUNCHES = [0]          # number of C-ABI calls issued (bench.py reports it)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    if isinstance(t, Mat):
        return t.ptr
    assert t.is_cuda, "music2dance_b200 kernels need CUDA tensors (no CPU fallback)"
    return t.data_ptr()


class Mat:
    """Channels-last activation view: element (b, l, c) at ptr + 4*(b*bs + l*ld + c)."""
    __slots__ = ("t", "ptr", "nb", "rows", "cols", "ld", "bs")

    def __init__(self, t, nb, rows, cols, ld=None, bs=None, offset=0):
        self.t = t
        self.ptr = t.data_ptr() + 4 * offset
        self.nb, self.rows, self.cols = nb, rows, cols
        self.ld = cols if ld is None else ld
        self.bs = rows * self.ld if bs is None else bs

    @staticmethod
    def of(t, nb, rows, cols):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        assert t.numel() == nb * rows * cols, (tuple(t.shape), nb, rows, cols)
        return Mat(t, nb, rows, cols)

    def cols_slice(self, c0, c1):
        m = Mat(self.t, self.nb, self.rows, c1 - c0, self.ld, self.bs)
        m.ptr = self.ptr + 4 * c0
        return m

    def batch_slice(self, b0, b1):
        m = Mat(self.t, b1 - b0, self.rows, self.cols, self.ld, self.bs)
        m.ptr = self.ptr + 4 * b0 * self.bs
        return m

    def flat_rows(self):
        """(nb, rows, C) viewed as (1, nb*rows, C); needs bs == rows*ld."""
        assert self.bs == self.rows * self.ld
        m = Mat(self.t, 1, self.nb * self.rows, self.cols, self.ld, self.nb * self.rows * self.ld)
        m.ptr = self.ptr
        return m

    def as_rows(self, nb, rows):
        """Re-interpret the batch/row split of a dense matrix (bs == rows*ld)."""
        assert self.bs == self.rows * self.ld and nb * rows == self.nb * self.rows
        m = Mat(self.t, nb, rows, self.cols, self.ld, rows * self.ld)
        m.ptr = self.ptr
        return m

    def flatten_cols(self):
        """(nb, rows, C) dense -> (nb, 1, rows*C) for full-length convolutions."""
        assert self.ld == self.cols and self.bs == self.rows * self.cols
        m = Mat(self.t, self.nb, 1, self.rows * self.cols)
        m.ptr = self.ptr
        return m

    @property
    def M(self):
        return self.nb * self.rows


def new_mat(nb, rows, cols, device, ld=None):
    ld = cols if ld is None else ld
    t = torch.empty(nb * rows * ld, dtype=torch.float32, device=device)
    return Mat(t, nb, rows, cols, ld)


# ---------------------------------------------------------------------------

def rowconv(x, w, y, *, T, Cc, N, sr=1, roff0=0, droff=1, w_ld=None, bias=None, act=0,
            mask=None, mask_mode=0, add=None, add_before_mask=False, y2=None, ws=None, win=None, w_tiled=None):
    """y = epi(rowconv(x, w)); see m2d_rowconv in include/m2d.h.  `win` = (T_frames,
    stride, pad, seq_len) switches on fused audio windowing (x = raw audio)."""
    a = RowConvArgs()
    if win is None:
        a.x, a.x_bs, a.x_ld, a.x_rows = x.ptr, x.bs, x.ld, x.rows
        a.win_T = 0
    else:
        a.x, a.x_bs, a.x_ld, a.x_rows = _p(x), 0, 1, win[4]
        a.win_T, a.win_stride, a.win_pad, a.win_seq_len = win[0], win[1], win[2], win[3]
    a.nb = y.nb
    a.w, a.w_ld = _p(w), (T * Cc if w_ld is None else w_ld)
    if w_tiled is not None:              # device address of the pre-split, pre-tiled weight copy (bulk-copy fed)
        a.w_tiled = w_tiled
    a.N, a.T, a.Cc = N, T, Cc
    a.sr, a.roff0, a.droff = sr, roff0, droff
    a.y, a.y_bs, a.y_ld, a.y_rows = y.ptr, y.bs, y.ld, y.rows
    a.y2 = None if y2 is None else y2.ptr
    if y2 is not None:
        assert (y2.bs, y2.ld) == (y.bs, y.ld)
    a.bias = _p(bias)
    a.act = act
    if mask is not None:
        a.mask, a.m_bs, a.m_ld, a.mask_mode = mask.ptr, mask.bs, mask.ld, mask_mode
    if add is not None:
        a.add, a.a_bs, a.a_ld, a.add_before_mask = add.ptr, add.bs, add.ld, int(add_before_mask)
    if ws is not None:
        a.ws, a.ws_floats = ws.data_ptr(), ws.numel()
    LAUNCHES[0] += 1
    call("m2d_rowconv", C.byref(a), _stream())


def wgrad(dy, x, dw, *, Cout, T, Cc, sr=1, roff0=0, droff=1, scale=1.0, beta=0.0, ws=None, win=None, packed=False):
    a = WgradArgs()
    a.dy, a.dy_bs, a.dy_ld, a.dy_rows = dy.ptr, dy.bs, dy.ld, dy.rows
    a.nb = dy.nb
    if win is None:
        a.x, a.x_bs, a.x_ld, a.x_rows = x.ptr, x.bs, x.ld, x.rows
        a.win_T = 0
    else:
        a.x, a.x_bs, a.x_ld, a.x_rows = _p(x), 0, 1, win[4]
        a.win_T, a.win_stride, a.win_pad, a.win_seq_len = win[0], win[1], win[2], win[3]
    a.Cout, a.T, a.Cc = Cout, T, Cc
    a.sr, a.roff0, a.droff = sr, roff0, droff
    a.dw = _p(dw)
    a.packed = int(packed)
    a.scale, a.beta = scale, beta
    a.ws, a.ws_floats = ws.data_ptr(), ws.numel()
    LAUNCHES[0] += 1
    call("m2d_wgrad", C.byref(a), _stream())


def pack_conv_fwd(w, wp, Cout, Cin, k):
    LAUNCHES[0] += 1
    call("m2d_pack_conv_fwd", _p(w), _p(wp), Cout, Cin, k, _stream())


def pack_conv_bwd(w, wd, Cout, Cin, k, stride):
    LAUNCHES[0] += 1
    call("m2d_pack_conv_bwd", _p(w), _p(wd), Cout, Cin, k, stride, _stream())


PACK_FWD, PACK_BWD, PACK_FULL_BWD, UNPACK_GRAD, PACK_BWD_MERGED = 0, 1, 2, 3, 4


def pack_table(entries, device):
    """entries: (w, dst|None, dst_tiled|None, Cout, Cin, k, stride, kind[, reserved]) -> device table for pack_batch."""
    import numpy as np
    dt = np.dtype([("w", "<u8"), ("dst", "<u8"), ("dst_tiled", "<u8"), ("Cout", "<i4"), ("Cin", "<i4"),
                   ("k", "<i4"), ("stride", "<i4"), ("kind", "<i4"), ("reserved", "<i4")])
    arr = np.zeros(len(entries), dtype=dt)
    ptr = lambda t: 0 if t is None else t.data_ptr()
    for i, e in enumerate(entries):
        w, dst, dtl, Cout, Cin, k, stride, kind = e[:8]
        arr[i] = (w.data_ptr(), ptr(dst), ptr(dtl), Cout, Cin, k, stride, kind, e[8] if len(e) > 8 else 0)
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return t, len(entries)


def tiled_floats(N, T, Cc):
    """Size (floats) of the w_tiled copy of an [N][T*Cc] weight operand (see include/m2d.h)."""
    R = 64 if N <= 64 else 128
    return -(-N // R) * T * -(-Cc // 32) * 2 * R * 32


def pack_batch(table, n):
    LAUNCHES[0] += 1
    call("m2d_pack_batch", _p(table), n, _stream())


def conv_dgrad_c1(dy, w, dx, *, nb, Lout, Cout, k, stride, pad, Lin):
    LAUNCHES[0] += 1
    call("m2d_conv_dgrad_c1", _p(dy), nb, Lout, Cout, _p(w), k, stride, pad, _p(dx), Lin, _stream())


def gru_forward(gi, w_hh, b_hh, h_out, ldh, save, B, T, H):
    LAUNCHES[0] += 1
    call("m2d_gru_forward", _p(gi), _p(w_hh), _p(b_hh), _p(h_out), ldh, _p(save), B, T, H, _stream())


def gru_backward(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H):
    LAUNCHES[0] += 1
    call("m2d_gru_backward", _p(dh_out), ldd, _p(h_out), ldh, _p(save), _p(w_hh), _p(dgi), _p(dgh),
         B, T, H, _stream())


def colstats(x, acc, groups=1):
    """acc[z*2C : (z+1)*2C] += column sums / sums of squares of row block z (x.M / groups rows each)."""
    assert x.M % groups == 0
    LAUNCHES[0] += 1
    call("m2d_colstats_groups", x.ptr, x.ld, x.M // groups, x.cols, groups, _p(acc), _stream())


def bn_apply(x, y, acc, gamma, beta, rm, rv, mr, act, momentum=0.1, eps=1e-5, groups=1):
    """Train-mode BatchNorm of `groups` consecutive row blocks, each with its own statistics; the running statistics
    advance block by block (groups separate forwards in one launch)."""
    assert x.M % groups == 0
    LAUNCHES[0] += 1
    call("m2d_bn_apply_groups", x.ptr, x.ld, None if y is None else y.ptr, 0 if y is None else y.ld, x.M // groups,
         x.cols, groups, _p(acc), _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps, _p(mr), act, _stream())


def bn_train(x, y, acc, gamma, beta, rm, rv, mr, act, momentum=0.1, eps=1e-5):
    """Train-mode BatchNorm, statistics + apply in one launch; acc: 2C + 1 zeroed doubles."""
    LAUNCHES[0] += 1
    call("m2d_bn_train", x.ptr, x.ld, None if y is None else y.ptr, 0 if y is None else y.ld, x.M, x.cols,
         _p(acc), _p(gamma), _p(beta), _p(rm), _p(rv), momentum, eps, _p(mr), act, _stream())


def bn_eval(x, y, gamma, beta, rm, rv, act, eps=1e-5):
    LAUNCHES[0] += 1
    call("m2d_bn_eval", x.ptr, x.ld, y.ptr, y.ld, x.M, x.cols, _p(gamma), _p(beta), _p(rm), _p(rv),
         eps, act, _stream())


def bn_bwd_reduce(dy, y, x, mr, act, acc):
    LAUNCHES[0] += 1
    call("m2d_bn_bwd_reduce", dy.ptr, dy.ld, y.ptr, y.ld, x.ptr, x.ld, x.M, x.cols, _p(mr), act,
         _p(acc), _stream())


def bn_bwd_apply(dy, y, x, dx, mr, gamma, act, acc, dgamma, dbeta):
    LAUNCHES[0] += 1
    call("m2d_bn_bwd_apply", dy.ptr, dy.ld, y.ptr, y.ld, x.ptr, x.ld, dx.ptr, dx.ld, x.M, x.cols,
         _p(mr), _p(gamma), act, _p(acc), _p(dgamma), _p(dbeta), _stream())


def colsum(x, out, acc, scale=1.0, beta=0.0):
    LAUNCHES[0] += 2
    call("m2d_colsum", x.ptr, x.ld, x.M, x.cols, _p(out), scale, beta, _p(acc), _stream())


def axpby(x, z, y, n, a=1.0, b=1.0):
    LAUNCHES[0] += 1
    call("m2d_axpby", _p(x), _p(z), _p(y), n, a, b, _stream())


def fill(y, n, v):
    LAUNCHES[0] += 1
    call("m2d_fill", _p(y), n, v, _stream())


def scale_rows(x, s, y, nb, per):
    LAUNCHES[0] += 1
    call("m2d_scale_rows", _p(x), _p(s), _p(y), nb, per, _stream())


def interp(real, fake, alpha, xi, nb, per):
    LAUNCHES[0] += 1
    call("m2d_interp", _p(real), _p(fake), _p(alpha), _p(xi), nb, per, _stream())


def rows_sumsq(x, nb, per, out):
    LAUNCHES[0] += 1
    call("m2d_rows_sumsq", _p(x), nb, per, _p(out), _stream())


def sum_(x, n, out):
    LAUNCHES[0] += 1
    call("m2d_sum", _p(x), n, _p(out), _stream())


def gp_finalize(ss0, ss1, B, gp, k0, k1, kscale=1.0):
    LAUNCHES[0] += 1
    call("m2d_gp_finalize", _p(ss0), _p(ss1), B, _p(gp), _p(k0), _p(k1), kscale, _stream())


def interp_stack3(real, fake, alpha, xi3, nb, per):
    """xi3 = [alpha*real + (1-alpha)*fake; real; fake]: the 3*nb stacked pose entries of a critic iteration."""
    LAUNCHES[0] += 1
    call("m2d_interp_stack3", _p(real), _p(fake), _p(alpha), _p(xi3), nb, per, _stream())


def colsum_table(entries, acc_doubles, device):
    """entries: (x Mat [1, M, C], out tensor, scale, beta) -> (device table, n, max C, fp64 scratch)."""
    import numpy as np
    dt = np.dtype([("x", "<u8"), ("out", "<u8"), ("M", "<i8"), ("acc_off", "<i8"), ("ld", "<i4"), ("C", "<i4"),
                   ("scale", "<f4"), ("beta", "<f4")])
    arr = np.zeros(len(entries), dtype=dt)
    off = 0
    for i, (x, out, scale, beta) in enumerate(entries):
        arr[i] = (x.ptr, out.data_ptr(), x.M, off, x.ld, x.cols, scale, beta)
        off += x.cols
    acc = torch.zeros(max(off, 1), dtype=torch.float64, device=device)
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    return t, len(entries), max(e[0].cols for e in entries), acc


def colsum_batch(tab):
    LAUNCHES[0] += 2
    call("m2d_colsum_batch", _p(tab[0]), tab[1], tab[2], _p(tab[3]), _stream())


def gp_finalize_lp(ss0, B, gp, k0):
    LAUNCHES[0] += 1
    call("m2d_gp_finalize_lp", _p(ss0), B, _p(gp), _p(k0), _stream())


def pose_losses(real, fake, dfake, B, T, Cn, beta, eta, accumulate, acc):
    LAUNCHES[0] += 1
    call("m2d_pose_losses", _p(real), _p(fake), _p(dfake), B, T, Cn, beta, eta, int(accumulate), _p(acc),
         _stream())


def jerkiness(x, B, T, Cn, acc):
    LAUNCHES[0] += 1
    call("m2d_jerkiness", _p(x), B, T, Cn, _p(acc), _stream())


def act_bwd(d, y, n, mode):
    LAUNCHES[0] += 1
    call("m2d_act_bwd", _p(d), _p(y), n, mode, _stream())


def mul3(a, b, c, out, alpha=1.0):
    """out = alpha * a * b * c on same-shaped Mats (strided rows)."""
    LAUNCHES[0] += 1
    call("m2d_mul3", a.ptr, a.ld, b.ptr, b.ld, c.ptr, c.ld, out.ptr, out.ld, a.M, a.cols, alpha, _stream())


def maxpool2(x, y, nb, Lin, Cn):
    LAUNCHES[0] += 1
    call("m2d_maxpool2", x.ptr, x.ld, y.ptr, y.ld, nb, Lin, Cn, _stream())


def maxpool2_bwd(x, dy, dx, nb, Lin, Cn, accumulate):
    LAUNCHES[0] += 1
    call("m2d_maxpool2_bwd", x.ptr, x.ld, dy.ptr, dy.ld, dx.ptr, dx.ld, nb, Lin, Cn, int(accumulate), _stream())


def upsample2(x, y, nb, Lin, Cn):
    LAUNCHES[0] += 1
    call("m2d_upsample2", x.ptr, x.ld, y.ptr, y.ld, nb, Lin, Cn, _stream())


def upsample2_bwd(dy, dx, nb, Lin, Cn, accumulate):
    LAUNCHES[0] += 1
    call("m2d_upsample2_bwd", dy.ptr, dy.ld, dx.ptr, dx.ld, nb, Lin, Cn, int(accumulate), _stream())


def copy2d(x, y, accumulate=False):
    LAUNCHES[0] += 1
    call("m2d_copy2d", x.ptr, x.ld, y.ptr, y.ld, x.M, x.cols, int(accumulate), _stream())


def fusion_mlp(x, w1, b1, w2, b2, u, d, dd=None, dh=None, dx=None):
    """u = relu(fc1(x)), d = fc2(u) on the rows of x [1,n,F]; with dd (n floats) also dh / dx (see m2d_fusion_mlp)."""
    H = w1.shape[0]
    assert x.nb == 1 and w1.shape[1] == x.cols and u.cols == H and u.ld == H and (dh is None or dh.ld == H)
    LAUNCHES[0] += 1
    call("m2d_fusion_mlp", x.ptr, x.ld, x.rows, x.cols, H, _p(w1), _p(b1), _p(w2), _p(b2),
         None if dd is None else dd.ptr, u.ptr, d.ptr, None if dh is None else dh.ptr,
         None if dx is None else dx.ptr, 0 if dx is None else dx.ld, _stream())


class Copy2dDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("y", C.c_void_p), ("M", C.c_longlong), ("ldx", C.c_int), ("ldy", C.c_int),
                ("C", C.c_int), ("accumulate", C.c_int)]


COPY2D_MAX = 4


def copy2d_batch(entries):
    """entries: up to 4 (x Mat, y Mat, accumulate) strided copies, ONE launch, applied in order."""
    assert 0 < len(entries) <= COPY2D_MAX
    arr = (Copy2dDesc * len(entries))()
    for i, (x, y, accumulate) in enumerate(entries):
        assert (x.M, x.cols) == (y.M, y.cols)
        arr[i] = Copy2dDesc(x.ptr, y.ptr, x.M, x.ld, y.ld, x.cols, int(accumulate))
    LAUNCHES[0] += 1
    call("m2d_copy2d_batch", C.cast(arr, C.c_void_p), len(entries), _stream())


def embed_rows(table, labels, y, B, T, n_classes, err=None):
    """y[(b*T+t), :E] = table[labels[b]] for a Mat `y` that is the label-column slice of a concatenated operand
    (conditional.py:19-21,45-46); labels int64 on the device."""
    assert labels.is_cuda and labels.dtype == torch.int64 and labels.is_contiguous() and labels.numel() == B
    LAUNCHES[0] += 1
    call("m2d_embed_rows", _p(table), labels.data_ptr(), y.ptr, y.ld, B, T, y.cols, n_classes, _p(err), _stream())


def embed_grad(dy, labels, dtable, B, T, n_classes, scale=1.0, beta=0.0):
    """dtable[c] (+)= scale * sum of the rows of `dy` (label-column slice) belonging to sequences labelled c."""
    assert labels.is_cuda and labels.dtype == torch.int64 and labels.numel() == B
    LAUNCHES[0] += 1
    call("m2d_embed_grad", dy.ptr, dy.ld, labels.data_ptr(), _p(dtable), B, T, dy.cols, n_classes, scale, beta,
         _stream())


def transpose_bcl(x, y, nb, R, Cn):
    """[b, R, C] -> [b, C, R] (dense)."""
    LAUNCHES[0] += 1
    call("m2d_transpose_bcl", _p(x), _p(y), nb, R, Cn, _stream())


def wgan_scalars(sums, gp, B, n_l1, n_tv, c0, c1, mode, out, d_real=None, d_fake=None):
    LAUNCHES[0] += 1
    call("m2d_wgan_scalars", _p(sums), _p(gp), B, n_l1, n_tv, c0, c1, mode, _p(out), _p(d_real), _p(d_fake), _stream())


def slice_audio(audio, out, nseq, A, nwin, W, stride, pad_left):
    LAUNCHES[0] += 1
    call("m2d_slice_audio", _p(audio), _p(out), nseq, A, nwin, W, stride, pad_left, _stream())


def crop_batch(poses, pose_off, music, music_off, seq, start, B, T, O, ratio, A, real, audio):
    LAUNCHES[0] += 1
    call("m2d_crop_batch", _p(poses), _p(pose_off), _p(music), _p(music_off), _p(seq), _p(start), B, T, O, ratio, A,
         _p(real), _p(audio), _stream())


def adam(p, g, m, v, n, step, lr, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0):
    LAUNCHES[0] += 2
    call("m2d_adam", _p(p), _p(g), _p(m), _p(v), n, _p(step), lr, b1, b2, eps, gscale, _stream())


def adam_pack(table, n, smem_floats, counters, lr, b1=0.9, b2=0.999, eps=1e-8, gscale=1.0, nbytes=0):
    """Adam + every weight re-layout of the table's parameters in ONE launch (m2d_adam_pack, include/m2d.h).
    nbytes: algorithmic bytes of the launch (engine.AdamPack.bytes), only read by bench.py's instrumentation."""
    LAUNCHES[0] += 1
    call("m2d_adam_pack", _p(table), n, smem_floats, _p(counters), lr, b1, b2, eps, gscale, _stream())


def nvl_allreduce(ptrs, mc, pads, rank, world, off, n, blocks, slot0, status):
    """ptrs / pads: ctypes arrays of `world` device pointers (this process's mappings of every rank's buffer / flags)."""
    LAUNCHES[0] += 1
    call("m2d_nvl_allreduce", ptrs, C.c_void_p(mc) if mc else None, pads, rank, world, off, n, blocks, slot0,
         _p(status), _stream())


def nvl_allreduce2(ptrs, mc, off, n, ptrs2, mc2, off2, n2, pads, rank, world, blocks, slot0, status):
    LAUNCHES[0] += 1
    call("m2d_nvl_allreduce2", ptrs, C.c_void_p(mc) if mc else None, off, n, ptrs2, C.c_void_p(mc2) if mc2 else None,
         off2, n2, pads, rank, world, blocks, slot0, _p(status), _stream())


GEMM_MODES = {"fp32": 0, "tf32": 1, "tf32bf16": 2, "tf32x3": 3}


def set_gemm_mode(mode):
    """Arithmetic of the GEMM family: 'fp32' (CUDA cores), 'tf32' (tcgen05, one TF32 product),
    'tf32x3' (tcgen05, 3xTF32 split, fp32-grade; the default), 'tf32bf16' (TF32 hi*hi + BF16 cross terms, opt-in;
    packed weights must be refreshed after switching to / from it).  See include/m2d.h."""
    call("m2d_set_gemm_mode", GEMM_MODES[mode] if isinstance(mode, str) else int(mode))


def get_gemm_mode():
    m = _lib.load().m2d_get_gemm_mode()
    return {v: k for k, v in GEMM_MODES.items()}[m]


def set_gru_impl(v):
    call("m2d_set_gru_impl", int(v))


class Trace:
    """Phase marks inside a (graph-captured) train step: `ops.mark(name)` enqueues a one-thread kernel that writes the
    device's global timer into a slot; disabled (the default) it is a no-op.  tools/step_timeline.py reads the slots
    after a replay, so the timeline is the one of the timed configuration, not of a profiler run."""

    def __init__(self, device, slots=4096):
        self.buf = torch.zeros(slots, dtype=torch.int64, device=device)
        self.names = []

    def mark(self, name):
        i = len(self.names)
        assert i < self.buf.numel()
        self.names.append((name, torch.cuda.current_stream().cuda_stream))
        call("m2d_timestamp", self.buf.data_ptr() + 8 * i, _stream())

    def read(self):
        t = self.buf[:len(self.names)].cpu().tolist()
        return [(n, s, v) for (n, s), v in zip(self.names, t)]


TRACE = [None]


def mark(name):
    tr = TRACE[0]
    if tr is not None:
        tr.mark(name)


def set_gru_forward_batch_group(bg):
    call("m2d_set_gru_forward_batch_group", int(bg))


def check_device(dev=0):
    call("m2d_check_device", dev)
    return _lib.load().m2d_version()
