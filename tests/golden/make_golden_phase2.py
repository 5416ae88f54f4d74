"""Golden values of the phase2 unconditional sequence WGAN-LP step (BASELINE.json configs[1]) from the UNMODIFIED
reference modules: phase2/archis/default.py SequenceGenerator / SequenceDiscriminator, losses.gradient_penalty
(is_seq=True, lp=True) and losses.tv_loss run the loop body of phase2/train.py:134-171 on CPU.
(build container only: needs /root/reference)"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase2_oracle as P2         # noqa: E402
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

B, SEED_STEP, SEED_DATA = 4, 31, 78


def put(out, prefix, t):
    d = O.tensor_digest(t)
    for k in ("sum", "l2", "maxabs"):
        out[f"{prefix}/{k}"] = np.float64(d[k])
    out[f"{prefix}/samples"] = d["samples"].numpy()


def main():
    _, losses, _ = R.import_reference()
    from phase2.archis.default import SequenceDiscriminator, SequenceGenerator
    cfg = P2.make_cfg()
    T, Oo = cfg["stick_length"], cfg["output_size"]
    out = {"B": B}
    for state in ("init", "perturbed"):
        torch.manual_seed(0)
        gen = SequenceGenerator(cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"], Oo,
                                cfg["nblocks_gen"], cfg["n_cells"], "cpu")
        critic = SequenceDiscriminator(Oo, cfg["channels"], T, init_ker=cfg["init_kernel"],
                                       n_blocks=cfg["nblocks_critic"], device="cpu")
        if state == "perturbed":
            for m in (gen, critic):
                sd = m.state_dict()
                O.perturb_params(sd)
                m.load_state_dict(sd)
        for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
            put(out, f"{state}/init/{k}", v)
        real_bt = P2.synthetic_poses(B, T, SEED_DATA)
        gen.train()
        torch.manual_seed(SEED_STEP)
        critic.zero_grad()
        noise = torch.randn(B, T, cfg["input_vector_size"])
        fake = gen(noise, [T] * B).view(B, T, Oo).permute(0, 2, 1).contiguous()
        real = real_bt.view(B, T, Oo).permute(0, 2, 1).contiguous()
        gp = losses.gradient_penalty(critic, B, real, fake, is_seq=True, lp=True, device=None)
        err_real = torch.mean(critic(real))
        err_fake = torch.mean(critic(fake.detach()))
        err = err_fake - err_real + cfg["gamma"] * gp
        err.backward(retain_graph=True)
        out[f"{state}/critic/loss_critic"] = np.float64(err.item())
        out[f"{state}/critic/gp"] = np.float64(gp.item())
        out[f"{state}/critic/w_dist"] = np.float64((err_fake - err_real).item())
        out[f"{state}/critic/fake"] = fake.detach().numpy()
        for k, p in critic.named_parameters():
            put(out, f"{state}/critic/grad/{k}", p.grad)
        gen.zero_grad()
        noise = torch.randn(B, T, cfg["input_vector_size"])
        fake = gen(noise, [T] * B).view(B, T, Oo).permute(0, 2, 1)
        err_real = torch.mean(critic(real))
        err_fake = torch.mean(critic(fake))
        err_tv = losses.tv_loss(fake)
        err_gen = err_real - err_fake + cfg["eta"] * err_tv
        err_gen.backward()
        out[f"{state}/gen/loss_gen"] = np.float64(err_gen.item())
        out[f"{state}/gen/tv"] = np.float64(err_tv.item())
        for k, p in gen.named_parameters():
            if p.grad is None:
                out[f"{state}/gen/nograd/{k}"] = np.int64(1)
            else:
                put(out, f"{state}/gen/grad/{k}", p.grad)
        for k, v in gen.state_dict().items():
            if "running" in k or "num_batches" in k:
                put(out, f"{state}/gen/buf/{k}", v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "phase2.npz"), **out)
    print({k: float(v) for k, v in out.items() if np.ndim(v) == 0 and "grad" not in k})


if __name__ == "__main__":
    main()
