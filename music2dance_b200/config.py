"""phase3 configuration constants and synthetic inputs of the reference shape for the product side
(bench.py, tools/): `phase3/configs/default.yaml:1-37` plus the constants train.py / utils.py derive from it
(`phase3/train.py:45-83`, `utils.py:50-58`).  Kept separate from oracle/ — product code never imports the checker."""
from __future__ import annotations

import torch

DEFAULT_CFG = dict(
    batch_size=7, window_size=0.2, seq_length=4.8, gamma=10.0, beta=1.0, eta=0.0,
    nblocks_gen=2, input_vector_size=250, latent_vector_size=250, n_cells=3,
    size=256, channels=128, output_size=69, lr_gen=2e-4, lr_critic=2e-4,
    n_critic_steps=8, code_size=100, noise_size=10, init_kernel=25,
    enc_type="default", ablated=False, activ="id",
    audio_rate=16000, video_rate=25,
)


def make_cfg(**over):
    cfg = dict(DEFAULT_CFG)
    cfg.update(over)
    cfg["stick_length"] = int(cfg["seq_length"] * cfg["video_rate"])          # utils.py:55 -> 120
    cfg["audio_length"] = int(cfg["seq_length"] * cfg["audio_rate"])          # utils.py:56 -> 76800
    cfg["cutting_stride"] = int(cfg["audio_rate"] / cfg["video_rate"])        # utils.py:57 -> 640
    cfg["audio_feat_samples"] = int(cfg["window_size"] * cfg["audio_rate"])   # train.py:82 -> 3200
    cfg["pad_samples"] = cfg["audio_feat_samples"] - cfg["cutting_stride"]    # train.py:83 -> 2560
    return cfg


def synthetic_batch(cfg, B, seed):
    """One iteration's inputs of the reference shape: real (B,T,23,3) ~ U[0,1) (MinMax-scaled poses),
    audio (B,A) ~ 0.3*U(-1,1), noise ~ N(0,1), alpha ~ U[0,1), noise_g ~ N(0,1), from a private generator."""
    g = torch.Generator().manual_seed(seed)
    T, A = cfg["stick_length"], cfg["audio_length"]
    real = torch.rand(B, T, 23, 3, generator=g)
    audio = (torch.rand(B, A, generator=g) * 2 - 1) * 0.3
    noise = torch.randn(B, T, cfg["noise_size"], generator=g)
    alpha = torch.rand(B, 1, generator=g)
    noise_g = torch.randn(B, T, cfg["noise_size"], generator=g)
    return real, audio, noise, alpha, noise_g
