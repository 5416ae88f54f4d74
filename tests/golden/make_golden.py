"""Generate golden fixtures from the UNMODIFIED reference modules.

    python tests/golden/make_golden.py [variant ...]    (build container only: needs /root/reference)

For every variant the reference's own SequenceGenerator / SequenceDiscriminator /
losses.gradient_penalty / utils.slice_audio_batch are executed on CPU (torch fp32)
for one critic iteration and one generator update from two states (seed-0 init,
and a deterministic RNG-free perturbation of it), and the outputs are stored as
small digests: scalars, the generated poses, per-tensor (sum, l2, max, 64 strided
samples) of every gradient / BN buffer, and the windowing of a short signal.
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

VARIANTS = {"default": {}, "wavegan": {"enc_type": "wavegan"}, "unet": {"enc_type": "unet"},
            "ablated": {"ablated": True}, "tanh": {"activ": "tanh"},
            # round 2: phase3/configs/tv.yaml, noise_enhanced.yaml, activ = 'relu' (same dict as tests/parity.py)
            "tv": {"eta": 50.0}, "noise_enhanced": {"noise_size": 100}, "relu": {"activ": "relu"}}
B = 2
ALPHA_SEED = 77
DATA_SEED = 1234


def put(out, prefix, d):
    for k in ("sum", "l2", "maxabs"):
        out[f"{prefix}/{k}"] = np.float64(d[k])
    out[f"{prefix}/samples"] = d["samples"].numpy()


MARGIN = 5e-7
REL_MARGIN = 1e-6
N_SEEDS = 48


def build_state(cfg, state):
    gen, critic = R.build_models(cfg, seed=0)
    if state == "perturbed":
        sg, sd = gen.state_dict(), critic.state_dict()
        O.perturb_params(sg)
        O.perturb_params(sd)
        gen.load_state_dict(sg)
        critic.load_state_dict(sd)
    return gen, critic


def kink_margin(cfg, gen, critic, seed):
    """Smallest |pre-activation| seen by the ReLUs of the critic's pose branch over the
    four pose inputs of one critic iteration + generator update (interpolate, real, fake,
    generator-step fake).  ReLU makes the gradients DISCONTINUOUS in the inputs: a
    pre-activation within fp32 noise (~1e-7) of zero flips its mask between two summation
    orders and moves every generator gradient by percents (observed: one unit at -3.4e-7
    -> 2.7 % on decoder.lastfc.weight).  These layers have only B*128*120 units, so the
    fixture seed is the best of N_SEEDS candidates and must keep a >= MARGIN (absolute) margin; the audio
    branch and the generator have ~1e6 units per layer, where a flip is both unavoidable
    and negligible.  The audio branch matters too: one unit of audio_d.l4 at 1.4e-8
    (5e-7 of that layer's rms) moved audio_d.l4.weight's gradient by 2 % between the CPU
    and the CUDA fp32 summation orders.  Those layers have ~2.4e6 units in total, so an
    absolute 1e-6 margin does not exist; they are screened RELATIVE to the layer's rms
    pre-activation (fp32 summation noise is ~1e-7 of it): min|pre|/rms >= REL_MARGIN."""
    _, _, utils = R.import_reference()
    lo = [float("inf")]
    hooks = []
    def hook_abs(mod, i, o):
        lo[0] = min(lo[0], float(o.detach().abs().min()))

    def hook_rel(mod, i, o):
        o = o.detach()
        rel = float(o.abs().min() / o.pow(2).mean().sqrt())
        lo[0] = min(lo[0], rel * (MARGIN / REL_MARGIN))      # normalised so that one threshold serves both
    for name, m in critic.named_modules():
        # every ReLU-fed layer: pose convs (absolute margin), audio_d.l1-l5 and fc1 (relative margin)
        if isinstance(m, (torch.nn.Conv1d, torch.nn.Linear)) and not name.endswith(("fconv", "l6", "fc2")):
            hooks.append(m.register_forward_hook(hook_abs if name.startswith("stick_d") else hook_rel))
    real, audio, noise, _, noise_g = O.synthetic_batch(cfg, B, seed)
    T, Oo = cfg["stick_length"], cfg["output_size"]
    sd = {k: v.clone() for k, v in gen.state_dict().items()}
    with torch.no_grad():
        sl = utils.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
        gen.train()
        fake = gen(sl, [T] * B, noise=noise).view(B, T, Oo).permute(0, 2, 1)
        fake_g = gen(sl, [T] * B, noise=noise_g).view(B, T, Oo).permute(0, 2, 1)
        rl = real.view(B, T, Oo).permute(0, 2, 1)
        torch.manual_seed(ALPHA_SEED)
        a = torch.rand(B, 1).view(B, 1, 1)
        for x in (rl, fake, fake_g, a * rl + (1 - a) * fake):
            if cfg["ablated"]:
                critic(x.contiguous())
            else:
                critic(x.contiguous(), audio.unsqueeze(1))
    gen.load_state_dict(sd)          # undo the BatchNorm running-stat updates
    for h in hooks:
        h.remove()
    return lo[0]


def one_state(cfg, state, out):
    gen, critic = build_state(cfg, state)
    cands = [(kink_margin(cfg, gen, critic, DATA_SEED + i), DATA_SEED + i) for i in range(N_SEEDS)]
    m, seed = max(cands)
    assert m > MARGIN, f"no seed with a ReLU margin above {MARGIN}: best {m:.2e}"
    out[f"{state}/data_seed"] = np.int64(seed)
    out[f"{state}/relu_margin"] = np.float64(m)
    for k, v in list(gen.state_dict().items()) + list(critic.state_dict().items()):
        put(out, f"{state}/init/{k}", O.tensor_digest(v))
    real, audio, noise, _, noise_g = O.synthetic_batch(cfg, B, seed)
    r = R.critic_iteration(gen, critic, cfg, real, audio, noise, ALPHA_SEED, None)
    for k in ("loss_critic", "gp", "w_dist", "err_real", "err_fake"):
        out[f"{state}/critic/{k}"] = np.float64(r[k])
    out[f"{state}/critic/fake"] = r["fake"].numpy()
    for k, g in r["grads"].items():
        put(out, f"{state}/critic/grad/{k}", O.tensor_digest(g))
    for k, v in gen.state_dict().items():
        if "running" in k or "num_batches" in k:
            put(out, f"{state}/critic/genbuf/{k}", O.tensor_digest(v))
    r = R.generator_update(gen, critic, cfg, real, audio, noise_g, None)
    for k in ("loss_gen", "l1", "tv", "err_real", "err_fake"):
        out[f"{state}/gen/{k}"] = np.float64(r[k])
    out[f"{state}/gen/fake"] = r["fake"].contiguous().numpy()
    for k, g in r["grads"].items():
        if g is None:
            out[f"{state}/gen/nograd/{k}"] = np.int8(1)
        else:
            put(out, f"{state}/gen/grad/{k}", O.tensor_digest(g))


def windowing(out):
    _, _, utils = R.import_reference()
    g = torch.Generator().manual_seed(5)
    for name, (n, win, stride) in {"a": (76800, 3200, 640), "b": (6400, 3200, 640),
                                   "c": (5000, 700, 160), "d": (3200, 3200, 640)}.items():
        x = torch.rand(2, n, generator=g)
        s = utils.slice_audio_batch(x, win, stride, win - stride)
        out[f"window/{name}/shape"] = np.array(s.shape)
        out[f"window/{name}/args"] = np.array([n, win, stride])
        put(out, f"window/{name}", O.tensor_digest(s, 256))
        s1 = utils.slice_audio_batch(x[0], win, stride, win - stride)
        assert torch.equal(s1, s[0])


if __name__ == "__main__":
    torch.set_num_threads(8)
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]                       # python tests/golden/make_golden.py [variant ...]
    for name, over in VARIANTS.items():
        if only and name not in only:
            continue
        cfg = O.make_cfg(**over)
        out = {}
        for state in ("init", "perturbed"):
            one_state(cfg, state, out)
        if name == "default":
            windowing(out)
        np.savez_compressed(os.path.join(here, f"phase3_{name}.npz"), **out)
        print(name, len(out), "entries", os.path.getsize(os.path.join(here, f"phase3_{name}.npz")) // 1024, "KiB")
