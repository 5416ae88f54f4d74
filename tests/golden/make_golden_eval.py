"""Golden values of the validation / evaluation path from the UNMODIFIED reference modules.

    python tests/golden/make_golden_eval.py      (build container only: needs /root/reference)

phase3/train.py:245-261 (eval-mode generator + mean L1 on a validation batch) and losses.jerkiness
(phase3/test.py:85-100) are executed with the reference's own SequenceGenerator / losses on CPU for the
three encoders, from the perturbed state (non-trivial BatchNorm running statistics), incl. a 750-frame
sequence for the default encoder (phase3/test.py:49)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

B, SEED = 3, 4321


def main():
    _, losses, utils = R.import_reference()
    out = {"B": B, "seed": SEED}
    for enc in ("default", "wavegan", "unet"):
        cfg = O.make_cfg(enc_type=enc)
        gen, _ = R.build_models(cfg, seed=0)
        sg = gen.state_dict()
        O.perturb_params(sg)
        gen.load_state_dict(sg)
        gen.eval()
        real_bt, audio, noise, _, _ = O.synthetic_batch(cfg, B, SEED)
        T, Oo = cfg["stick_length"], cfg["output_size"]
        with torch.no_grad():
            sl = utils.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
            real = real_bt.view(B, T, Oo).permute(0, 2, 1)
            fake = gen(sl, [T] * B, noise=noise).view(B, T, Oo).permute(0, 2, 1)
            out[f"{enc}/l1_val"] = np.float64(torch.nn.L1Loss(reduction="mean")(real, fake).item())
            out[f"{enc}/fake"] = fake.numpy()
            out[f"{enc}/jerk_fake"] = np.float64(losses.jerkiness(fake).item())
            out[f"{enc}/jerk_real"] = np.float64(losses.jerkiness(real).item())
            # test.py form: one long (1, 69, B*T) sequence
            out[f"{enc}/jerk_fake_flat"] = np.float64(losses.jerkiness(fake.permute(0, 2, 1).reshape(1, -1, Oo).permute(0, 2, 1)).item())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "phase3_eval.npz"), **out)
    print({k: float(v) for k, v in out.items() if np.ndim(v) == 0})


if __name__ == "__main__":
    main()
