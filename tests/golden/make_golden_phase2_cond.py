"""Golden values of the phase2 CONDITIONAL sequence WGAN step (BASELINE.json configs[2]) from the UNMODIFIED reference
modules: phase2/archis/conditional.py SequenceGenerator / SequenceDiscriminator (train mode: both Dropout layers
active, masks drawn from the global CPU generator), losses.gradient_penalty (is_seq=True, lp=True) on
``lambda x: critic(x, labels)`` and losses.tv_loss, run through the loop body of phase2/train.py:134-171 with the
labels threaded through (phase2/train_conditional.py does not match these modules; see oracle/phase2_cond_oracle.py).
(build container only: needs /root/reference)"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase2_cond_oracle as PC    # noqa: E402
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

B, SEED_STEP, SEED_DATA, SEED_LABEL = 4, 47, 91, 5


def put(out, prefix, t):
    d = O.tensor_digest(t)
    for k in ("sum", "l2", "maxabs"):
        out[f"{prefix}/{k}"] = np.float64(d[k])
    out[f"{prefix}/samples"] = d["samples"].numpy()


def main():
    _, losses, _ = R.import_reference()
    from phase2.archis.conditional import SequenceDiscriminator, SequenceGenerator
    cfg = PC.make_cfg()
    T, Oo = cfg["stick_length"], cfg["output_size"]
    out = {"B": B}
    for state in ("init", "perturbed"):
        torch.manual_seed(0)
        gen = SequenceGenerator(cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"], Oo,
                                cfg["nblocks_gen"], cfg["n_cells"])
        critic = SequenceDiscriminator(Oo, cfg["channels"], T, init_ker=cfg["init_kernel"],
                                       n_blocks=cfg["nblocks_critic"])
        if state == "perturbed":
            for m in (gen, critic):
                sd = m.state_dict()
                O.perturb_params(sd)
                m.load_state_dict(sd)
        for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in critic.state_dict().items()]:
            put(out, f"{state}/init/{k}", v)
        real_bt = PC.synthetic_poses(B, T, SEED_DATA)
        labels = PC.synthetic_labels(B, SEED_LABEL)
        out["labels"] = labels.numpy()
        gen.train()
        critic.train()
        torch.manual_seed(SEED_STEP)
        critic.zero_grad()
        noise = torch.randn(B, T, cfg["input_vector_size"])
        fake = gen(noise, labels).view(B, T, Oo).permute(0, 2, 1).contiguous()
        real = real_bt.view(B, T, Oo).permute(0, 2, 1).contiguous()
        gp = losses.gradient_penalty(lambda x: critic(x, labels), B, real, fake, is_seq=True, lp=True, device=None)
        err_real = torch.mean(critic(real, labels))
        err_fake = torch.mean(critic(fake.detach(), labels))
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
        fake = gen(noise, labels).view(B, T, Oo).permute(0, 2, 1)
        err_real = torch.mean(critic(real, labels))
        err_fake = torch.mean(critic(fake, labels))
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
        # eval mode (no dropout, running statistics): scores and poses on fixed inputs
        gen.eval()
        critic.eval()
        with torch.no_grad():
            noise = torch.randn(B, T, cfg["input_vector_size"], generator=torch.Generator().manual_seed(3))
            out[f"{state}/eval/noise_seed"] = np.int64(3)
            out[f"{state}/eval/fake"] = gen(noise, labels).numpy()
            out[f"{state}/eval/score"] = critic(real, labels).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "phase2_cond.npz"), **out)
    print({k: float(v) for k, v in out.items() if np.ndim(v) == 0 and "grad" not in k and "init" not in k})


if __name__ == "__main__":
    main()
