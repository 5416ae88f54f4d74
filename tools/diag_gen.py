"""Diagnostic (GPU): generator parameter-gradient error vs the fp64 oracle (smooth upstream)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import phase3_oracle as O
from tests.parity import VARIANTS
from tests.test_parity_gpu import build, DEV, oracle_params
from music2dance_b200.utils import slice_audio_batch

for variant in sys.argv[1:] or ["wavegan"]:
    cfg = O.make_cfg(**VARIANTS[variant])
    gen, _ = build(cfg, "perturbed")
    B, T = 3, 120
    g = torch.Generator().manual_seed(21)
    audio = (torch.rand(B, cfg["audio_length"], generator=g) - 0.5) * 0.6
    noise = torch.randn(B, T, cfg["noise_size"], generator=g)
    up = torch.randn(B * T, cfg["output_size"], generator=g)
    sl = O.slice_audio_batch(audio, 3200, 640, 2560)
    G = oracle_params(gen)
    names = O.trainable_names(G)
    res = {}
    for dt in (torch.float32, torch.float64):
        P = O._leaf({k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in G.items()})
        out = O.generator_forward(P, cfg, sl.to(dt), noise.to(dt), train=True)
        gl = torch.autograd.grad((out * up.to(dt)).sum(), [P[k] for k in names], allow_unused=True)
        res[dt] = (out.detach(), dict(zip(names, gl)))
    gen.train()
    out = gen(slice_audio_batch(audio.to(DEV), 3200, 640, 2560), [T] * B, noise=noise.to(DEV))
    (out * up.to(DEV)).sum().backward()
    o64, g64 = res[torch.float64]
    o32, g32 = res[torch.float32]
    print("==", variant, "out err ours", float((out.detach().cpu().double() - o64).abs().max() / o64.abs().max()),
          "ref32", float((o32.double() - o64).abs().max() / o64.abs().max()))
    skip = set(O.pre_bn_bias_names(G))
    got = dict(gen.named_parameters())
    for k in names:
        if g64[k] is None or k in skip:
            continue
        s = float(g64[k].abs().max().clamp_min(1e-12))
        eo = float((got[k].grad.cpu().double() - g64[k]).abs().max()) / s
        er = float((g32[k].double() - g64[k]).abs().max()) / s
        l2 = float((got[k].grad.cpu().double() - g64[k]).norm() / g64[k].norm())
        print(f"  {k:48s} ours {eo:.2e} (l2 {l2:.2e})  ref32 {er:.2e}  max|g| {s:.2e}")
