"""Diagnostic (GPU): per-parameter critic-gradient error of the drop-in path vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import phase3_oracle as O
from tests.parity import VARIANTS, load_golden, B_GOLD, ALPHA_SEED
from tests.test_parity_gpu import build, DEV
from music2dance_b200.losses import gradient_penalty
from music2dance_b200.utils import slice_audio_batch

for variant in sys.argv[1:] or ["unet", "default"]:
    cfg = O.make_cfg(**VARIANTS[variant])
    gold = load_golden(variant)
    gen, critic = build(cfg, "init")
    B, T, Oo = B_GOLD, 120, 69
    real_bt, audio, noise, _, noise_g = O.synthetic_batch(cfg, B, int(gold["init/data_seed"]))
    G = {k: v.cpu().clone() for k, v in gen.state_dict().items()}
    D = {k: v.cpu().clone() for k, v in critic.state_dict().items()}
    gen.train()
    slices = slice_audio_batch(audio, 3200, 640, 2560)
    real = real_bt.to(DEV).view(B, T, Oo).permute(0, 2, 1).contiguous()
    aud = audio.to(DEV).unsqueeze(1)
    fake = gen(slices, [T] * B, noise=noise.to(DEV)).view(B, T, Oo).permute(0, 2, 1).contiguous().detach()
    # oracle on the SAME fake
    Dl = O._leaf(D)
    fk = fake.cpu()
    rl = real.cpu()
    ac = audio.unsqueeze(1)
    names = O.trainable_names(D)
    def ograds(expr):
        gl = torch.autograd.grad(expr, [Dl[k] for k in names], allow_unused=True, retain_graph=True)
        return dict(zip(names, gl))
    er = O.critic_forward(Dl, cfg, rl, ac).mean()
    ef = O.critic_forward(Dl, cfg, fk, ac).mean()
    go_r, go_f = ograds(er), ograds(ef)
    def ours(x):
        critic.zero_grad()
        critic(x, aud).mean().backward()
        return {k: p.grad.detach().cpu().clone() for k, p in critic.named_parameters()}
    gr, gf = ours(real), ours(fake)
    print("==", variant)
    for k in names:
        for nm, a, b in (("real", gr[k], go_r[k]), ("fake", gf[k], go_f[k])):
            e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
            if e > 2e-4:
                print(f"  {k:32s} {nm} err {e:.3e}  max|ref| {float(b.abs().max()):.3e}")
    d = {k: (gf[k] - gr[k]) for k in names}
    do = {k: (go_f[k] - go_r[k]) for k in names}
    for k in names:
        e = float((d[k] - do[k]).abs().max() / do[k].abs().max().clamp_min(1e-12))
        if e > 2e-4:
            print(f"  {k:32s} diff err {e:.3e}")
    # GP part
    torch.manual_seed(ALPHA_SEED)
    alpha = torch.rand(B, 1)
    gp_o, _, _ = O.gradient_penalty(Dl, cfg, rl, fk, ac, alpha)
    go_gp = ograds(gp_o)
    critic.zero_grad()
    torch.manual_seed(ALPHA_SEED)
    gp = gradient_penalty(critic, B, real, fake, aud, is_seq=True, lp=False, device=DEV)
    gp.backward()
    print("  gp", float(gp), float(gp_o))
    for k, p in critic.named_parameters():
        b = go_gp[k]
        if b is None:
            continue
        a = p.grad.cpu() if p.grad is not None else torch.zeros_like(b)
        e = float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
        if e > 2e-4:
            print(f"  {k:32s} GP err {e:.3e}  max|ref| {float(b.abs().max()):.3e}")
