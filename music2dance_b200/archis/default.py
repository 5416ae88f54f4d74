"""Drop-in for the reference's ``phase3/archis/default.py``.

Same class names, constructor arguments, ``forward`` signatures, public attributes
and ``state_dict`` keys / shapes / order, and the same initial
weights under the same ``torch.manual_seed`` (constructor RNG consumption matches the
reference: default torch init in creation order, then ``initialize_weights``).

The ``torch.nn`` layers created here are *parameter containers only*: ``forward``
never calls them.  All arithmetic runs in hand-written sm_100a kernels
(libm2d_b200.so) through ``music2dance_b200.nets``; there is no PyTorch / CPU fallback
and forward raises if the inputs are not on a CUDA device.

Autograd contract: first-order ``backward()`` through generator and critic is
supported (gradients w.r.t. inputs and parameters).  Double backward through the critic
(``autograd.grad(..., create_graph=True)``) is NOT routed through autograd — use
``music2dance_b200.losses.gradient_penalty`` (same signature as the reference's), which
evaluates penalty and its weight gradients with fused kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..engine import engine_of
from ..ops import Mat
from ..utils import initialize_weights


class _Holder(nn.Module):
    """Structural node of the parameter tree (attribute names define state_dict keys)."""

    def forward(self, *a, **k):
        raise RuntimeError("parameter container: the computation runs in music2dance_b200 CUDA kernels "
                           "through the owning SequenceGenerator / SequenceDiscriminator")


_NODE_CLASSES = {}


def _node(cls_name):
    """Structural node that prints like the reference's module of that name (`str(module)` goes into model_gen.txt /
    model_critic.txt, phase3/train.py:173-178): a _Holder subclass named after the reference class."""
    cls = _NODE_CLASSES.get(cls_name)
    if cls is None:
        cls = _NODE_CLASSES[cls_name] = type(cls_name, (_Holder,), {})
    return cls()


def _node_class(path, enc_type=None):
    """Reference class of the structural node at dotted `path` (default.py: who owns which attribute)."""
    import re
    table = [(r"audio_enc$", "AudioEncoder"),
             (r"audio_enc\.model$", {"default": "DefaultAudioEncoder", "unet": "UNetAudioEncoder",
                                     "wavegan": "WaveGANAudioEncoder"}.get(enc_type, "_Holder")),
             (r"audio_enc\.model\.(conv_layers|activations)$", "ModuleList"),
             (r"audio_enc\.model\.activations\.\d+$", "Sequential"),
             (r"audio_enc\.model\.ublock$", "UBlock"),
             (r"audio_enc\.model\.ublock\.convblock\d$", "BasisConvBlock"),
             (r"(audio_rnn|noise_gen)$", "NoiseGen"),
             (r"decoder$", "FrameDecoder"), (r"decoder\.blocks$", "Sequential"), (r"decoder\.blocks\.\d+$", "LinearBlock"),
             (r"stick_d$", "StickDiscriminator"), (r"stick_d\.blocks$", "Sequential"),
             (r"stick_d\.blocks\.\d+$", "TemporalBlock"), (r"audio_d$", "AudioDiscriminator")]
    for pat, name in table:
        if re.match(pat, path):
            return name
    return "_Holder"


def _attach(root, dotted, leaf, enc_type=None):
    node = root
    parts = dotted.split(".")
    for i, p in enumerate(parts[:-1]):
        if p not in node._modules:
            name = _node_class(".".join(parts[:i + 1]), enc_type)
            node.add_module(p, _Holder() if name == "_Holder" else _node(name))
        node = node._modules[p]
    node.add_module(parts[-1], leaf)


def _code_activ(activ):
    """default.py:71-76,98-103,133-138,308-313,336-341: the module the reference applies to a code."""
    if activ == 'id':
        return nn.Identity()
    if activ == 'relu':
        return nn.ReLU(True)
    if activ == 'tanh':
        return nn.Tanh()
    raise ValueError(f"unknown activ {activ!r} (id | relu | tanh)")


def _conv(ci, co, k, s=1, p=0):
    return nn.Conv1d(ci, co, k, stride=s, padding=p)


def _encoder_layers(enc_type, f, out, activ='id'):
    """(dotted name under audio_enc.model, layer) in the reference's creation order; parameter-free layers (ReLU,
    LeakyReLU, Identity / Tanh, MaxPool1d, Upsample) are listed too so that the module tree prints like the
    reference's — they draw no random numbers and own no state."""
    L = []
    if enc_type == "default":                      # default.py:59-76
        c = [1, f, 2 * f, 4 * f, 8 * f, 16 * f, 32 * f]
        L.append(("conv_layers.0", _conv(1, f, 250, 50, 124)))
        for i in range(1, 6):
            L.append((f"conv_layers.{i}", _conv(c[i], c[i + 1], 4, 2, 1)))
        L.append(("conv_layers.6", _conv(c[6], out, 2)))
        for i in range(6):
            L.append((f"activations.{i}.0", nn.BatchNorm1d(c[i + 1])))
            L.append((f"activations.{i}.1", nn.ReLU(True)))
        L.append(("activations.6", _code_activ(activ)))
    elif enc_type == "wavegan":                    # default.py:114-135
        c = [1, f, 2 * f, 4 * f, 8 * f]
        for i in range(1, 5):
            L.append((f"l{i}", _conv(c[i - 1], c[i], 25, 4)))
            L.append((f"bn{i}", nn.BatchNorm1d(c[i])))
        L.append(("l5", _conv(c[4], out, 5)))
        L.append(("relu", nn.ReLU(True)))
        L.append(("activ", _code_activ(activ)))
    elif enc_type == "unet":                       # default.py:85-104,213-239
        L.append(("conv_layers.0", _conv(1, f, 160, 4, 79)))
        L.append(("conv_layers.1", _conv(f, 2 * f, 4, 2, 1)))
        L.append(("conv_layers.2", _conv(2 * f, 4 * f, 4, 2, 1)))
        for i, ch in enumerate((f, 2 * f, 4 * f)):
            L.append((f"activations.{i}.0", nn.BatchNorm1d(ch)))
            L.append((f"activations.{i}.1", nn.LeakyReLU(0.2)))
        ch = 4 * f
        for i in range(1, 8):
            L.append((f"ublock.convblock{i}.conv", _conv(ch if i <= 4 else 2 * ch, ch, 3, 1, 1)))
            L.append((f"ublock.convblock{i}.bn", nn.BatchNorm1d(ch)))
            L.append((f"ublock.convblock{i}.relu", nn.LeakyReLU(0.2)))
        L.append(("ublock.downsample", nn.MaxPool1d(2, 2)))
        L.append(("ublock.upsample", nn.Upsample(scale_factor=2, mode="linear", align_corners=False)))
        L.append(("fc", _conv(ch, out, 200)))
        L.append(("activ", _code_activ(activ)))
    else:
        raise ValueError(f"unknown enc_type {enc_type!r} (default | unet | wavegan)")
    return L


class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, x, noise, *params):
        eng = owner._engine()
        eng.ensure_packed()
        B, T, W = x.shape
        net = eng.net
        with torch.cuda.device(x.device):
            fake = net.forward(None, noise.contiguous().float(), B, T, train=owner.training,
                               slices=x.contiguous().float())
            eng.fid += 1
            ctx.owner, ctx.fid, ctx.n = owner, eng.fid, B * T
            return fake.t[:B * T * net.O].view(B * T, net.O).clone()

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.owner._engine()
        if eng.fid != ctx.fid:
            raise RuntimeError("SequenceGenerator: backward through a forward whose activations were "
                               "overwritten by a later forward (keep one live graph per generator)")
        with torch.cuda.device(dout.device):
            d = dout.contiguous().float()
            eng.net.backward(Mat.of(d, 1, ctx.n, eng.net.O))
            return (None, None, None, *eng.grads_in_param_order())


class SequenceGenerator(nn.Module):
    """Audio windows (B,T,W) -> poses (B*T, output_size).  Reference: default.py:6-42."""

    def __init__(self, window_size, input_size, latent_size, size, output_size, noise_size, n_blocks,
                 n_cells=1, enc_type="default", activ='id', device="cpu"):
        super().__init__()
        self.window_size = window_size
        self.input_size = input_size
        self.latent_size = latent_size
        self.size = size
        self.noise_size = noise_size
        self.output_size = output_size
        self.device = device
        self.n_blocks, self.n_cells, self.enc_type, self.activ = n_blocks, n_cells, enc_type, activ
        for name, layer in _encoder_layers(enc_type, 32, input_size, activ):
            _attach(self, "audio_enc.model." + name, layer, enc_type)
        _attach(self, "audio_rnn.rnn", nn.GRU(input_size, latent_size - noise_size, n_cells, batch_first=True))
        _attach(self, "noise_gen.rnn", nn.GRU(noise_size, noise_size, 1, batch_first=True))
        _attach(self, "decoder.fc1", nn.Linear(latent_size, size))
        _attach(self, "decoder.bn1", nn.BatchNorm1d(size, eps=1e-5, momentum=0.1))
        _attach(self, "decoder.relu", nn.ReLU(inplace=True))
        if n_blocks == 0:
            self._modules["decoder"].add_module("blocks", _node("Sequential"))
        for b in range(n_blocks):
            q = f"decoder.blocks.{b}."
            _attach(self, q + "fc1", nn.Linear(size, size))
            _attach(self, q + "fc2", nn.Linear(size, size))
            _attach(self, q + "bn1", nn.BatchNorm1d(size, eps=1e-5, momentum=0.1))
            _attach(self, q + "bn2", nn.BatchNorm1d(size, eps=1e-5, momentum=0.1))
            _attach(self, q + "relu", nn.ReLU(inplace=True))
        _attach(self, "decoder.lastfc", nn.Linear(size, output_size))
        dec = self._modules["decoder"]
        dec.latent_size, dec.size, dec.output_size, dec.nblocks = latent_size, size, output_size, n_blocks
        initialize_weights(self)
        self.to(device)

    def _cfg(self):
        return dict(enc_type=self.enc_type, activ=self.activ, audio_feat_samples=self.window_size,
                    input_vector_size=self.input_size, latent_vector_size=self.latent_size,
                    noise_size=self.noise_size, n_cells=self.n_cells, size=self.size,
                    output_size=self.output_size, nblocks_gen=self.n_blocks,
                    cutting_stride=getattr(self, "cutting_stride", 640),
                    pad_samples=getattr(self, "pad_samples", self.window_size - 640))

    def _engine(self):
        return engine_of(self, "gen", self._cfg)

    def forward(self, x, lengths, noise=None):
        """x (B, T, window_size); `lengths` as in the reference (all equal to T, Q15);
        noise (B, T, noise_size) or None (drawn on the CPU generator like the reference, Q3)."""
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.SequenceGenerator needs CUDA inputs (no CPU fallback)")
        B, T = x.size(0), x.size(1)
        if any(int(l) != T for l in lengths):
            raise NotImplementedError("ragged `lengths` are not used by phase3/train.py (all sequences have "
                                      "stick_length frames); only equal lengths are implemented")
        if noise is None:
            noise = torch.randn([B, T, self.noise_size]).to(x.device)
        self._engine()                      # flatten parameters before autograd sees them
        return _GeneratorFn.apply(self, x, noise, *self.parameters())


class _CriticFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, x, c, *params):
        eng = owner._engine()
        eng.ensure_packed()
        D = eng.net
        B = x.shape[0]
        with torch.cuda.device(x.device):
            slot, gen = eng.next_slot()
            tag = f"m{slot}"
            X = D.wk.mat(f"{tag}:X", B, D.T, D.O)
            xt = x.transpose(1, 2)
            if xt.is_contiguous() and xt.dtype == torch.float32:       # Q8: permuted view of (B,T,O)
                ops.copy2d(Mat.of(xt, 1, B * D.T, D.O), X.flat_rows())
            else:
                ops.transpose_bcl(x.contiguous().float(), X, B, D.O, D.T)
            from ..wgan import critic_forward
            aud = None if D.ablated else c.contiguous().float().view(B, -1)
            D.wk.acc_reset()
            fw = critic_forward(D, X, aud, B, B, tag)
            ctx.owner, ctx.fw, ctx.slot, ctx.gen, ctx.B, ctx.tag = owner, fw, slot, gen, B, tag
            ctx.need_x = x.requires_grad
            ctx.need_c = (c is not None) and c.requires_grad and not D.ablated
            ctx.cshape = None if c is None else c.shape
            return fw["d"].t[:B].view(B, 1).clone()

    @staticmethod
    def backward(ctx, dd):
        if torch.is_grad_enabled():
            # autograd.grad(..., create_graph=True): the stock losses.gradient_penalty (losses.py:40-44) asks for a
            # differentiable backward.  The kernels evaluate the penalty's double backward as a tangent pass instead.
            raise RuntimeError(
                "music2dance_b200 critic: backward with create_graph=True (double backward through the critic) is not "
                "routed through autograd — use music2dance_b200.losses.gradient_penalty (same signature as the "
                "reference's losses.gradient_penalty), which returns the penalty with its weight gradients attached")
        eng = ctx.owner._engine()
        if eng.slot_gen.get(ctx.slot) != ctx.gen:
            raise RuntimeError("critic: saved activations were overwritten (more than 6 live critic graphs)")
        D, fw, B, tag = eng.net, ctx.fw, ctx.B, ctx.tag + "b"
        wk = D.wk
        with torch.cuda.device(dd.device):
            wk.acc_reset()
            ddm = Mat.of(dd.contiguous().float().view(-1), 1, B, 1)
            D.fc2.wgrad(ddm, fw["u"], wk.scratch, acc=wk.acc_slot(1))
            dh, dsa = D.fusion_bwd(ddm, fw["u"], B, tag)
            D.fc1.wgrad(dh, fw["sa"], wk.scratch, acc=wk.acc_slot(128))
            d_s = wk.mat(f"{tag}:d_s", 1, B, D.code)
            ops.copy2d(dsa.cols_slice(0, D.code), d_s)
            dX = wk.mat(f"{tag}:dX", B, D.T, D.O) if ctx.need_x else None
            D.pose_bwd(fw["svp"], d_s, B, tag, wgrads=True, dX=dX)
            gx = gc = None
            if dX is not None:
                gx = torch.empty(B, D.O, D.T, device=dd.device)
                ops.transpose_bcl(dX, gx, B, D.T, D.O)
            if not D.ablated:
                d_a = wk.mat(f"{tag}:d_a", 1, B, D.code)
                ops.copy2d(dsa.cols_slice(D.code, D.F), d_a)
                if ctx.need_c:
                    gc = torch.empty(ctx.cshape, device=dd.device)
                D.audio_bwd(fw["sva"], d_a, B, tag, wgrads=True, dX=gc)
            return (None, gx, gc, *eng.grads_in_param_order())


def _critic_cfg(self):
    k6 = self._modules["audio_d"]._modules["l6"].kernel_size[0] if "audio_d" in self._modules else 0
    return dict(ablated="audio_d" not in self._modules, activ=self.activ, output_size=self.channels_in,
                channels=self.channels_h, code_size=self.output_code, stick_length=self.seqlen,
                audio_length=k6 * 4 ** 5)


def _build_stick_d(root, channels_in, channels_h, output_code, seqlen, init_ker, n_blocks=2, activ='id'):
    _attach(root, "stick_d.conv1", _conv(channels_in, channels_h, init_ker, 1, int((init_ker - 1) / 2)))
    for b in range(n_blocks):
        _attach(root, f"stick_d.blocks.{b}.conv1", _conv(channels_h, channels_h, 7, 1, 3))
        _attach(root, f"stick_d.blocks.{b}.conv2", _conv(channels_h, channels_h, 7, 1, 3))
        _attach(root, f"stick_d.blocks.{b}.relu", nn.ReLU(inplace=True))
    _attach(root, "stick_d.fconv", _conv(channels_h, output_code, seqlen))
    _attach(root, "stick_d.relu", nn.ReLU(inplace=True))
    _attach(root, "stick_d.activ", _code_activ(activ))


class SequenceDiscriminator(nn.Module):
    """Critic: pose branch + raw-audio branch + fusion MLP.  Reference: default.py:249-270.
    forward(x (B,69,T), c (B,1,A)) -> (B,1)."""

    def __init__(self, channels_in, channels_h, output_code, seqlen, init_ker=9, activ='id', device="cpu"):
        super().__init__()
        self.channels_in, self.channels_h, self.output_code, self.seqlen = channels_in, channels_h, output_code, seqlen
        self.activ = activ
        _build_stick_d(self, channels_in, channels_h, output_code, seqlen, init_ker, activ=activ)
        ch = [1, 32, 64, 128, 256, 512]
        for i in range(1, 6):                                    # default.py:298-302
            _attach(self, f"audio_d.l{i}", _conv(ch[i - 1], ch[i], 25, 4, 11))
        _attach(self, "audio_d.l6", _conv(512, output_code, 75))
        _attach(self, "audio_d.relu", nn.ReLU(True))
        _attach(self, "audio_d.activ", _code_activ(activ))
        self.fc1 = nn.Linear(2 * output_code, 128)
        self.fc2 = nn.Linear(128, 1)
        self.relu = nn.ReLU(True)
        initialize_weights(self)
        self.to(device)

    _cfg = _critic_cfg

    def _engine(self):
        return engine_of(self, "critic", self._cfg)

    def forward(self, x, c):
        if not (x.is_cuda and c.is_cuda):
            raise RuntimeError("music2dance_b200.SequenceDiscriminator needs CUDA inputs (no CPU fallback)")
        self._engine()
        return _CriticFn.apply(self, x, c, *self.parameters())


class AblatedSequenceDiscriminator(nn.Module):
    """Pose-only critic.  Reference: default.py:273-291 (Q9: `init_ker` is accepted but NOT
    forwarded, so the first convolution always has kernel 9 / padding 4)."""

    def __init__(self, channels_in, channels_h, output_code, seqlen, init_ker=9, activ='id', device="cpu"):
        super().__init__()
        self.channels_in, self.channels_h, self.output_code, self.seqlen = channels_in, channels_h, output_code, seqlen
        self.activ = activ
        _build_stick_d(self, channels_in, channels_h, output_code, seqlen, 9, activ=activ)
        self.fc1 = nn.Linear(output_code, 128)
        self.fc2 = nn.Linear(128, 1)
        self.relu = nn.ReLU(True)
        initialize_weights(self)
        self.to(device)

    _cfg = _critic_cfg

    def _engine(self):
        return engine_of(self, "critic", self._cfg)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("music2dance_b200.AblatedSequenceDiscriminator needs CUDA inputs (no CPU fallback)")
        self._engine()
        return _CriticFn.apply(self, x, None, *self.parameters())
