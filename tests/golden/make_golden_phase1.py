"""Golden values of the phase1 stick-figure WGAN-GP step (BASELINE.json configs[0]) from the UNMODIFIED
reference modules: phase1/archis/residual.py Generator / Discriminator and losses.gradient_penalty run the
loop body of phase1/train_wgan-gp.py:80-106 on CPU.      (build container only: needs /root/reference)"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase1_oracle as P1         # noqa: E402
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

SEED_INIT, SEED_STEP, SEED_DATA = 0, 2024, 77


def put(out, prefix, t):
    d = O.tensor_digest(t)
    for k in ("sum", "l2", "maxabs"):
        out[f"{prefix}/{k}"] = np.float64(d[k])
    out[f"{prefix}/samples"] = d["samples"].numpy()


def main():
    _, losses, _ = R.import_reference()
    from phase1.archis.residual import Discriminator, Generator
    out = {}
    for name, over in (("b2l50s256", {}), ("b1l10s32", dict(nblocks_gen=1, nblocks_critic=1, latent_vector_size=10, size=32))):
        cfg = P1.make_cfg(**over)
        B = cfg["batch_size"]
        torch.manual_seed(SEED_INIT)
        gen = Generator(cfg["latent_vector_size"], cfg["size"], cfg["output_size"], cfg["nblocks_gen"])
        critic = Discriminator(cfg["output_size"], cfg["size"], cfg["nblocks_critic"])
        for m in (gen, critic):
            sd = m.state_dict()
            O.perturb_params(sd)
            m.load_state_dict(sd)
        for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
            put(out, f"{name}/init/{k}", v)
        real = P1.synthetic_poses(B, SEED_DATA)
        gen.train()
        critic.train()
        torch.manual_seed(SEED_STEP)
        # train_wgan-gp.py:80-94
        critic.zero_grad()
        noise = torch.randn(B, cfg["latent_vector_size"])
        fake = gen(noise)
        gp = losses.gradient_penalty(critic, B, real, fake, device=None)
        err_real = torch.mean(critic(real))
        err_fake = torch.mean(critic(fake.detach()))
        err = err_fake - err_real + cfg["gamma"] * gp
        err.backward()
        out[f"{name}/critic/loss_critic"] = np.float64(err.item())
        out[f"{name}/critic/gp"] = np.float64(gp.item())
        out[f"{name}/critic/w_dist"] = np.float64((err_fake - err_real).item())
        out[f"{name}/critic/fake"] = fake.detach().numpy()
        for k, p in critic.named_parameters():
            if p.grad is None:
                out[f"{name}/critic/nograd/{k}"] = np.int64(1)
            else:
                put(out, f"{name}/critic/grad/{k}", p.grad)
        # train_wgan-gp.py:97-106 (no optimiser step in between: fixtures check one phase at a time)
        gen.zero_grad()
        noise = torch.randn(B, cfg["latent_vector_size"])
        fake = gen(noise)
        err_real = torch.mean(critic(real))
        err_fake = torch.mean(critic(fake))
        err_gen = err_real - err_fake
        err_gen.backward()
        out[f"{name}/gen/loss_gen"] = np.float64(err_gen.item())
        for k, p in gen.named_parameters():
            if p.grad is None:
                out[f"{name}/gen/nograd/{k}"] = np.int64(1)
            else:
                put(out, f"{name}/gen/grad/{k}", p.grad)
        for k, v in gen.state_dict().items():
            if "running" in k or "num_batches" in k:
                put(out, f"{name}/gen/buf/{k}", v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "phase1.npz"), **out)
    print({k: float(v) for k, v in out.items() if np.ndim(v) == 0 and "/critic/" in k and "grad" not in k})


if __name__ == "__main__":
    main()
