"""Input pipeline of the phase3 train loop on a device-resident dataset.

Mirrors what the reference does between its dataset and the first line of the loop body:
``SequenceDataset.__getitem__`` / ``get_positions`` (utils.py:91-101,245-248), ``collate_fn``
(utils.py:128-144), ``DataLoader(train_dataset, batch_size, sampler=WeightedRandomSampler(...),
collate_fn=collate_fn)`` (phase3/train.py:133-158) and the ``.to(device)`` copies (:191-193).

The whole dataset (61 sequences: ~1 GB of audio at most) lives in HBM; a batch is one gather kernel
(``m2d_crop_batch``) driven by B sequence indices and B start frames.  The random draws stay on the host
and consume the SAME generators in the SAME order as the reference (torch default generator for the
DataLoader base seed and the sampler's multinomial, numpy's global generator for the crop positions), so
seeding ``torch`` / ``numpy`` reproduces the reference's batches bit for bit while the per-step
host->device traffic drops from 2.2 MB (B = 7) to 2·B integers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class DeviceSequenceDataset:
    """Device-resident ``SequenceDataset(name_dict, config, resume=True, withaudio=True)``.

    `data` is the dict the reference builds in phase3/train.py:147-154: ``sequences`` (list of
    (L_i, 23, 3) arrays), ``musics`` (list of 1-D arrays), ``labels``, ``dirs``."""

    def __init__(self, data, config, device="cuda"):
        self.aud_rate, self.vid_rate = config["audio_rate"], config["video_rate"]
        self.seq_length = config["seq_length"]
        self.stick_length = int(config["seq_length"] * self.vid_rate)          # utils.py:55
        self.audio_length = int(config["seq_length"] * self.aud_rate)          # utils.py:56
        self.ratio = int(config["audio_rate"] / config["video_rate"])          # utils.py:57
        self.labels, self.dirs = data["labels"], data["dirs"]
        seqs = [np.asarray(s) for s in data["sequences"]]
        mus = [np.asarray(m) for m in data["musics"]]
        self.lengths = [len(s) for s in seqs]
        self.O = int(np.prod(seqs[0].shape[1:]))
        self.joint_shape = tuple(seqs[0].shape[1:])
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceSequenceDataset needs a CUDA device (no CPU fallback)")
        self.device = dev
        # collate_fn writes the crops into a float32 tensor and __getitem__ calls .float() on the music:
        # both are element-wise casts, done once here
        po = np.cumsum([0] + [s.shape[0] * self.O for s in seqs])
        mo = np.cumsum([0] + [m.shape[0] for m in mus])
        self.music_lengths = [m.shape[0] for m in mus]
        self.poses = torch.from_numpy(np.concatenate([s.reshape(-1) for s in seqs]).astype(np.float32)).to(dev)
        self.music = torch.from_numpy(np.concatenate(mus).astype(np.float32)).to(dev)
        self.pose_off = torch.from_numpy(po[:-1].astype(np.int64)).to(dev)
        self.music_off = torch.from_numpy(mo[:-1].astype(np.int64)).to(dev)

    def __len__(self):
        return len(self.lengths)

    def positions(self, idx):
        """get_positions (utils.py:245-248): one draw from numpy's global generator."""
        s = np.random.randint(0, self.lengths[idx] - self.stick_length)
        return s, s + self.stick_length

    def crop(self, seq_idx, starts, real=None, audio=None):
        """Batch of crops: real (B, T, *joint_shape), audio (B, A) on the device."""
        B, T, A = len(seq_idx), self.stick_length, self.audio_length
        for i, s in zip(seq_idx, starts):
            if not (0 <= s and s + T <= self.lengths[i] and s * self.ratio + A <= self.music_lengths[i]):
                raise IndexError(f"crop [{s}, {s + T}) outside sequence {i} "
                                 f"({self.lengths[i]} frames, {self.music_lengths[i]} samples)")
        if real is None:
            real = torch.empty((B, T) + self.joint_shape, dtype=torch.float32, device=self.device)
        if audio is None:
            audio = torch.empty(B, A, dtype=torch.float32, device=self.device)
        ctl = torch.tensor([list(seq_idx), list(starts)], dtype=torch.int32).to(self.device, non_blocking=True)
        with torch.cuda.device(self.device):
            ops.crop_batch(self.poses, self.pose_off, self.music, self.music_off, ctl[0], ctl[1], B, T, self.O,
                           self.ratio, A, real, audio)
        return real, audio


class Phase3Loader:
    """``DataLoader(dataset, batch_size, sampler=WeightedRandomSampler(weights, len(weights)),
    collate_fn=collate_fn)`` (phase3/train.py:133-158) over a DeviceSequenceDataset: iterating yields
    ``(padded_seqs, lengths, musics, labels, dirs)`` like collate_fn, the two tensors already on the device."""

    def __init__(self, dataset, batch_size, weights=None):
        self.ds, self.batch_size = dataset, batch_size
        self.weights = None if weights is None else torch.as_tensor(weights, dtype=torch.double)

    def __len__(self):
        return -(-len(self.ds) // self.batch_size)

    def _indices(self):
        # DataLoader draws its base seed from the default generator when the iterator is created,
        # then WeightedRandomSampler.__iter__ draws the whole epoch with one multinomial
        torch.empty((), dtype=torch.int64).random_()
        if self.weights is None:
            return torch.randperm(len(self.ds)).tolist()
        return torch.multinomial(self.weights, len(self.weights), True).tolist()

    def __iter__(self):
        idx = self._indices()
        for b0 in range(0, len(idx), self.batch_size):
            bi = idx[b0:b0 + self.batch_size]
            starts = [self.ds.positions(i)[0] for i in bi]          # __getitem__ order = sampler order
            # collate_fn sorts by length, descending, stable: every crop has stick_length frames -> order kept
            real, audio = self.ds.crop(bi, starts)
            labels = torch.stack([torch.from_numpy(np.asarray(self.ds.labels[i])) for i in bi])
            yield real, [self.ds.stick_length] * len(bi), audio, labels, tuple(self.ds.dirs[i] for i in bi)
