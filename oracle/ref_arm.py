"""Reference arm of bench.py: the reference's OWN modules (oracle/_ref, staged by oracle/build_ref.py; falls back to
/root/reference in the build container) driven through the loop body of phase3/train.py:186-237, on the host CPU
(`--impl reference`) or on cuda:0 through PyTorch eager + cuDNN/cuBLAS (`torch_eager_gpu`, the bar a user of the
reference gets on the same B200 today).

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by music2dance_b200/).  The loop body is restated line by line
around the stock modules because train.py itself cannot run (dataset absent, yaml.load without Loader):
  :188 zero_grad  :189-190 slice_audio_batch on the CPU  :191-194 .to(device)  :195 gen forward (noise drawn inside on
  the CPU generator)  :196-199 views  :204-205 gradient_penalty  :210-211 critic real / fake  :212-216 backward + Adam
  :218-219 gate  :221-237 generator update.
"""
from __future__ import annotations

import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_root():
    cands = [os.path.join(HERE, "_ref"), os.environ.get("M2D_REFERENCE", "/root/reference")]
    for c in cands:
        if os.path.exists(os.path.join(c, "phase3", "archis", "default.py")):
            return c
    return None


def import_reference():
    root = ref_root()
    if root is None:
        raise RuntimeError("reference modules not staged (run oracle/build_ref.py where /root/reference exists)")
    if "librosa" not in sys.modules:
        sys.modules["librosa"] = types.ModuleType("librosa")      # only used by the dataset loader (utils.py:185)
    for p in (root, os.path.join(root, "phase3")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import archis.default as archis
    import losses
    import utils
    return root, archis, losses, utils


class ReferenceStep:
    """One reference train step (n_critic critic iterations + 1 generator update) on `device`."""

    def __init__(self, cfg, B, device="cpu", tf32=False, seed=0):
        self.root, archis, self.losses, self.utils = import_reference()
        self.cfg, self.B, self.device = cfg, B, device
        if device != "cpu":
            torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
            torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.manual_seed(seed)                                    # train.py:35
        self.gen = archis.SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"],
                                            cfg["latent_vector_size"], cfg["size"], cfg["output_size"],
                                            cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"],
                                            cfg["enc_type"], cfg["activ"], device)
        cls = archis.AblatedSequenceDiscriminator if cfg["ablated"] else archis.SequenceDiscriminator
        self.critic = cls(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                          init_ker=cfg["init_kernel"], activ=cfg["activ"], device=device)
        self.optim_gen = torch.optim.Adam(self.gen.parameters(), lr=cfg["lr_gen"])           # train.py:102-103
        self.optim_critic = torch.optim.Adam(self.critic.parameters(), lr=cfg["lr_critic"])
        self.criterion = torch.nn.L1Loss(reduction="mean")                                   # train.py:170
        self.logs = {}

    def step(self, batches):
        """batches: n_critic tuples (real (B,T,23,3), audio (B,A)) of CPU tensors (what the DataLoader yields)."""
        cfg, B, dev = self.cfg, self.B, self.device
        T, O = cfg["stick_length"], cfg["output_size"]
        gen, critic = self.gen, self.critic
        gen.train()
        n = len(batches)
        for it, (real, audio) in enumerate(batches, 1):
            self.optim_critic.zero_grad()
            audio_slices = self.utils.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"],
                                                        cfg["pad_samples"])
            real = real.to(dev)
            audio = audio.to(dev)
            audio_slices = audio_slices.to(dev)
            audio = audio.unsqueeze(1)
            fake = gen(audio_slices, [T] * B)
            fake = fake.view(B, T, O).permute(0, 2, 1).contiguous()
            real = real.view(B, T, O).permute(0, 2, 1).contiguous()
            if cfg["ablated"]:
                gp = self.losses.gradient_penalty(critic, B, real, fake, is_seq=True, lp=False, device=dev)
                err_real = torch.mean(critic(real))
                err_fake = torch.mean(critic(fake.detach()))
            else:
                gp = self.losses.gradient_penalty(critic, B, real, fake, audio, is_seq=True, lp=False, device=dev)
                err_real = torch.mean(critic(real, audio))
                err_fake = torch.mean(critic(fake.detach(), audio))
            err_critic = err_fake - err_real + cfg["gamma"] * gp
            loss_critic = err_critic.item()
            err_critic.backward(retain_graph=True)
            self.optim_critic.step()
            if it % n:
                continue
            self.optim_gen.zero_grad()
            fake = gen(audio_slices, [T] * B)
            fake = fake.view(B, T, O).permute(0, 2, 1)
            err_l1 = self.criterion(real, fake)
            if cfg["ablated"]:
                err_real = torch.mean(critic(real))
                err_fake = torch.mean(critic(fake))
            else:
                err_real = torch.mean(critic(real, audio))
                err_fake = torch.mean(critic(fake, audio))
            err_tv = self.losses.tv_loss(fake)
            err_gen = err_real - err_fake + cfg["beta"] * err_l1 + cfg["eta"] * err_tv
            loss_gen = err_gen.item()
            err_gen.backward()
            self.optim_gen.step()
            self.logs = {"loss_critic": loss_critic, "gp": gp.item(), "loss_gen": loss_gen, "l1": err_l1.item()}
        if dev != "cpu":
            torch.cuda.synchronize()


def time_steps(cfg, B, device, steps, warmup, budget_s, tf32=False, threads=None):
    """Median seconds per train step over `steps` timed steps after `warmup` untimed ones, both cut down (never below
    3 timed / 1 warm-up) so that the whole run stays inside `budget_s`.  Returns a dict."""
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(threads or cores)
    g = torch.Generator().manual_seed(4321)
    nc, T, A = cfg["n_critic_steps"], cfg["stick_length"], cfg["audio_length"]

    def batches():
        return [(torch.rand(B, T, 23, 3, generator=g), (torch.rand(B, A, generator=g) * 2 - 1) * 0.3)
                for _ in range(nc)]

    rs = ReferenceStep(cfg, B, device, tf32=tf32)
    t0 = time.perf_counter()
    rs.step(batches())                                  # first warm-up step (thread pools, cuDNN autotune, allocator)
    t_first = time.perf_counter() - t0
    done_w = 1
    t1 = None
    if warmup > 1:
        t0 = time.perf_counter()
        rs.step(batches())
        t1 = time.perf_counter() - t0
        done_w = 2
    est = t1 if t1 is not None else t_first
    left = budget_s - t_first - (t1 or 0.0)
    k = int(max(3, min(steps, left / max(est, 1e-6))))
    w_more = int(max(0, min(warmup - done_w, (left - k * est) / max(est, 1e-6))))
    for _ in range(w_more):
        rs.step(batches())
    done_w += w_more
    ts = []
    for _ in range(k):
        bs = batches()
        t0 = time.perf_counter()
        rs.step(bs)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"s_per_step": med, "steps_timed": k, "warmup_done": done_w, "p10_s": ts[max(0, int(0.1 * k))],
            "p90_s": ts[min(k - 1, int(0.9 * k))], "cores": cores if device == "cpu" else 0,
            "threads": torch.get_num_threads() if device == "cpu" else 0,
            "source": ("oracle/_ref (unmodified reference modules staged by oracle/build_ref.py)"
                       if rs.root.endswith("_ref") else rs.root),
            "logs": rs.logs, "torch": torch.__version__, "tf32": bool(tf32)}
