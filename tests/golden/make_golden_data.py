"""Golden vectors of the phase3 input pipeline from the UNMODIFIED reference classes.

    python tests/golden/make_golden_data.py      (build container only: needs /root/reference)

The reference's own SequenceDataset(resume=True, withaudio=True), collate_fn, WeightedRandomSampler and
DataLoader (phase3/train.py:133-158) iterate two epochs over the seeded synthetic dataset of
oracle/data_oracle.py; every batch is stored in full (poses, audio, labels are small at 9 sequences)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import data_oracle as DO          # noqa: E402
from oracle import reference_harness as R      # noqa: E402

SEED_T, SEED_N, BATCH = 321, 654, 4


def main():
    _, _, utils = R.import_reference()
    from torch.utils.data import DataLoader, WeightedRandomSampler
    data, cfg = DO.synthetic_dataset(), DO.dataset_config()
    ds = utils.SequenceDataset(data, cfg, resume=True, withaudio=True)
    w = DO.class_weights(data["labels"])
    sampler = WeightedRandomSampler(w, len(w))
    loader = DataLoader(ds, batch_size=BATCH, sampler=sampler, collate_fn=utils.collate_fn)
    torch.manual_seed(SEED_T)
    np.random.seed(SEED_N)
    out = {"seed_torch": SEED_T, "seed_numpy": SEED_N, "batch": BATCH}
    n = 0
    for ep in range(2):
        for real, lengths, audio, labels, dirs in loader:
            out[f"b{n}/real"] = real.numpy()
            out[f"b{n}/audio_digest"] = np.asarray([float(audio.double().sum()), float(audio.double().abs().max())])
            out[f"b{n}/audio_head"] = audio[:, :64].numpy()
            out[f"b{n}/audio_tail"] = audio[:, -64:].numpy()
            out[f"b{n}/labels"] = labels.numpy()
            out[f"b{n}/lengths"] = np.asarray(lengths)
            out[f"b{n}/dirs"] = np.asarray(list(dirs))
            n += 1
    out["n_batches"] = n
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "phase3_data.npz"), **out)
    print("batches", n, "bytes", os.path.getsize(os.path.join(ROOT, "tests", "golden", "phase3_data.npz")))


if __name__ == "__main__":
    main()
