"""TEST INFRASTRUCTURE — CPU restatement of the reference's phase3 input pipeline (only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package).

Follows: SequenceDataset.__getitem__ (utils.py:91-101), get_positions (utils.py:245-248), collate_fn
(utils.py:128-144) and the loader of phase3/train.py:133-158 (WeightedRandomSampler + DataLoader with
num_workers = 0).  Pinned against the reference's own classes by tests/golden/make_golden_data.py
(fixture tests/golden/phase3_data.npz)."""
from __future__ import annotations

import numpy as np
import torch


def dataset_config():
    """configs/default.yaml `dataset:` block (seq_length 4.8 s, 25 fps, 16 kHz)."""
    return dict(audio_rate=16000, video_rate=25, seq_length=4.8, feat_size=69)


def synthetic_dataset(n=9, seed=5):
    """Small seeded stand-in for the 61-sequence dataset: ragged lengths, float64 poses (what
    MinMaxScaler.transform returns, utils.py:80-86), 4 dance-type labels."""
    g = np.random.RandomState(seed)
    seqs, mus, labels, dirs = [], [], [], []
    for i in range(n):
        L = int(g.randint(130, 220))
        seqs.append(g.rand(L, 23, 3))
        mus.append(((g.rand(L * 640 + int(g.randint(0, 500))) * 2 - 1) * 0.3).astype(np.float32))
        labels.append(int(g.randint(0, 4)))
        dirs.append(f"DANCE_{'WCRT'[labels[-1]]}_{i}")
    return dict(sequences=seqs, musics=mus, labels=np.asarray(labels), dirs=dirs)


def getitem(data, idx, T, ratio, A):
    """utils.py:91-101 (withaudio=True); get_positions draws from numpy's global generator."""
    seq = data["sequences"][idx]
    s = np.random.randint(0, len(seq) - T)
    s_a = s * ratio
    return (torch.from_numpy(seq[s:s + T]), torch.from_numpy(data["musics"][idx][s_a:s_a + A]).float(),
            torch.from_numpy(np.asarray(data["labels"][idx])), data["dirs"][idx]), s


def collate(batch):
    """utils.py:128-144."""
    batch.sort(key=lambda x: len(x[0]), reverse=True)
    sequences, musics, labels, dirs = zip(*batch)
    musics = torch.stack(musics)
    labels = torch.stack(labels)
    lengths = [len(seq) for seq in sequences]
    padded = torch.zeros(len(sequences), max(lengths), 23, 3)
    for i, seq in enumerate(sequences):
        padded[i, :lengths[i]] = seq[:lengths[i]]
    return padded, lengths, musics, labels, dirs


def class_weights(labels):
    """phase3/train.py:133-138."""
    count = np.unique(labels, return_counts=True)[1]
    return (1.0 / count)[labels]


def epoch(data, cfg, batch_size, weights):
    """One pass of the train loader; yields (collated batch, sequence indices, start frames)."""
    T = int(cfg["seq_length"] * cfg["video_rate"])
    A = int(cfg["seq_length"] * cfg["audio_rate"])
    ratio = int(cfg["audio_rate"] / cfg["video_rate"])
    torch.empty((), dtype=torch.int64).random_()                       # DataLoader base seed
    idx = torch.multinomial(torch.as_tensor(weights, dtype=torch.double), len(weights), True).tolist()
    for b0 in range(0, len(idx), batch_size):
        bi = idx[b0:b0 + batch_size]
        items, starts = [], []
        for i in bi:
            it, s = getitem(data, i, T, ratio, A)
            items.append(it)
            starts.append(s)
        yield collate(items), bi, starts
