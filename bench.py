#!/usr/bin/env python
"""bench.py — phase3 WGAN-GP train steps/sec (seq = 120) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--enc default]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (reference phase3/train.py:186-237 until the `continue`
at :219 falls through): n_critic = 8 critic iterations, each on its own fresh batch (generator
forward with train-mode BatchNorm, gradient penalty, Wasserstein terms, backward, Adam), then
one generator update (forward, L1 + critic terms + TV, backward, Adam).  Workload =
BASELINE.json configs[3] (default.yaml: default audio encoder, 4.8 s audio / 120 frames,
batch 7 per GPU).  Data: synthetic tensors of the reference shape; random-init weights.

Printed (rank 0, ONE JSON line):
  value   train steps/s summed over ranks (weak scaling: every rank runs batch-B steps on its own
          shard; gradients are all-reduced over NCCL), inputs already resident in HBM
  e2e     the same through the public API with HOST (pinned) input buffers: per step one
          host->device copy of the step's inputs and one device->host read of its scalars
  roofline   dominant kernel family, algorithmic FLOPs / CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline   the oracle port of the reference path timed on this box's host cores (N = 1)

`--impl reference` times the CPU implementation of the same path (the oracle port: the
reference is Python/PyTorch and /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "phase3 WGAN-GP train steps/sec (seq=120)"
UNIT = "train steps/s"
L2_BYTES = 126 << 20
TRAFFIC_FILE = "r02_traffic.json"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=7, help="sequences per GPU per iteration (default.yaml: 7)")
    ap.add_argument("--enc", default="default", choices=["default", "wavegan", "unet"])
    ap.add_argument("--gemm", default="tf32x3", choices=["fp32", "tf32", "tf32bf16", "tf32x3"],
                    help="arithmetic of the GEMM family (include/m2d.h): tcgen05 3xTF32 split (default, fp32-grade), "
                         "tcgen05 single-pass TF32, or CUDA-core fp32")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-device-dataset", action="store_true",
                    help="skip the e2e_device_dataset measurement (input pipeline on a device-resident dataset)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0,
                    help="wall-clock bound of the reference arm (steps / warm-up are cut down to fit, never below 3 / 1)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: host CPU (the reference arm) or cuda:0 through PyTorch eager + cuDNN/cuBLAS "
                         "(used by the torch_eager_gpu leg)")
    ap.add_argument("--ref-tf32", action="store_true", help="--ref-device cuda: allow TF32 in cuDNN / cuBLAS")
    ap.add_argument("--no-eager-gpu", action="store_true", help="skip the torch_eager_gpu leg (reference modules on cuda:0)")
    ap.add_argument("--no-throughput-regime", action="store_true",
                    help="skip the throughput_regime sub-record (a second, short run at --regime-batch per GPU)")
    ap.add_argument("--regime-batch", type=int, default=512,
                    help="per-GPU batch of the throughput_regime sub-record (SURVEY 8d: the reference batch 7 is far too "
                         "small to load a B200; 512 is its throughput configuration)")
    ap.add_argument("--collective", default=None, choices=["nvl", "nccl"],
                    help="N > 1: gradient all-reduce by the hand-written NVLink peer-memory kernel inside ONE step graph "
                         "(nvl) or by NCCL between per-iteration graphs (nccl); default: nvl when available")
    ap.add_argument("--sub", action="store_true", help=argparse.SUPPRESS)      # child run: no nested legs
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def workload_name(args, cfg):
    return (f"phase3 audio-conditioned WGAN-GP, {args.enc} audio encoder, 4.8 s synthetic audio/120 frames, "
            f"batch {args.batch}/GPU, n_critic {cfg['n_critic_steps']}")


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's own modules (oracle/_ref, staged by oracle/build_ref.py) through the loop body of
# phase3/train.py:186-237 — PyTorch CPU eager on the host cores, or eager + cuDNN/cuBLAS on cuda:0
# ----------------------------------------------------------------------------------------------

def reference_time(cfg, B, device, steps, warmup, budget_s, tf32=False):
    """Returns (seconds per train step, description of the sample, cores, kind, detail dict)."""
    from oracle import ref_arm
    if ref_arm.ref_root() is not None:
        r = ref_arm.time_steps(cfg, B, device, steps, warmup, budget_s, tf32=tf32)
        sample = (f"{r['steps_timed']} full train steps ({cfg['n_critic_steps']} critic iterations + 1 generator update "
                  f"each) at batch {B} after {r['warmup_done']} warm-up step(s), median; fresh synthetic batches, "
                  f"windowing on the CPU and H2D copies inside the step as in train.py:189-193")
        return r["s_per_step"], sample, r["cores"], "reference", r
    # the staged reference is missing (oracle/build_ref.py never ran): the oracle port, full steps as well
    import torch
    from oracle import phase3_oracle as O
    assert device == "cpu", "the oracle port is a CPU restatement"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    G, D = O.init_generator_params(cfg), O.init_critic_params(cfg)
    ad, ag = O.AdamState(D, cfg["lr_critic"]), O.AdamState(G, cfg["lr_gen"])
    t0 = time.perf_counter()
    O.train_step(G, D, cfg, B, 0, ag, ad)
    est = time.perf_counter() - t0
    k = int(max(3, min(steps, (budget_s - est) / max(est, 1e-6))))
    ts = []
    for i in range(k):
        t0 = time.perf_counter()
        O.train_step(G, D, cfg, B, 1 + i, ag, ad)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], f"{k} full train steps of the oracle port at batch {B} after 1 warm-up step, median", cores, "port", \
        {"steps_timed": k, "warmup_done": 1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from music2dance_b200 import config as O
    cfg = O.make_cfg(enc_type=args.enc)
    t0 = time.perf_counter()
    t_step, sample, cores, kind, det = reference_time(cfg, args.batch, args.ref_device, args.steps, args.warmup,
                                                      args.cpu_budget_s, tf32=args.ref_tf32)
    v = 1.0 / t_step
    where = ("host CPU (PyTorch eager fp32)" if args.ref_device == "cpu" else
             "cuda:0 (PyTorch eager, cuDNN/cuBLAS, " + ("TF32 allowed" if args.ref_tf32 else "fp32, TF32 off") + ")")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": det["steps_timed"], "steps_requested": args.steps, "warmup": det["warmup_done"],
            "warmup_requested": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if not args.ref_tf32 else "tf32", "data": "synthetic",
            "config": {"workload": workload_name(args, cfg), "device": where, "global_batch": args.batch,
                       "note": "one process at the per-GPU batch: the reference has no multi-GPU path; compare with the "
                               "N = 1 line only"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
            "detail": {k: det[k] for k in ("p10_s", "p90_s", "threads", "source", "torch", "logs") if k in det}}
    print(json.dumps(line), flush=True)


def child_line(extra, timeout_s):
    """Run this script again in a child process (own CUDA context, memory released afterwards) and return its JSON line."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__)] + extra
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as e:                                                  # noqa: BLE001
        return {"error": repr(e)[:400]}


def measure_tf32_peak(torch, dev, seconds=3.0):
    """Dense TF32 tensor-pipe peak of this GPU the way MEASURED_PEAKS.json measures bf16: cuBLAS 8192^3 fp32 matmul
    with TF32 allowed — best of 10 (burst) and back to back for `seconds` (sustained, under the power cap)."""
    n = 8192
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        fl = 2.0 * n ** 3
        return {"tf32_tflops": fl / (best * 1e-3) / 1e12, "tf32_tflops_sustained": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": f"torch.matmul fp32 {n}^3 with allow_tf32 (cuBLAS): best of 10 (burst) and {reps} back to back (sustained)"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.sm)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# per-kernel-family instrumentation (roofline)
# ----------------------------------------------------------------------------------------------

class KernelTimer:
    """Wraps the C-ABI wrappers in music2dance_b200.ops with CUDA events recorded on the launching
    stream (torch's current stream == the stream handed to the C ABI) and the algorithmic
    FLOPs / bytes of each launch."""

    def __init__(self, ops, torch):
        self.ops, self.torch = ops, torch
        self.rec = []            # (family, flops, bytes, ev0, ev1)
        self.saved = {}

    def _wrap(self, name, fam_fn):
        orig = getattr(self.ops, name)
        self.saved[name] = orig
        torch = self.torch

        def wrapped(*a, **k):
            fam, flops, nbytes = fam_fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **k)
            e1.record()
            self.rec.append((fam, flops, nbytes, e0, e1))
            return r
        setattr(self.ops, name, wrapped)

    def install(self):
        def f_rowconv(x, w, y, *, T, Cc, N, **k):
            M = y.nb * y.rows
            fam = "rowconv_c1" if Cc == 1 else "rowconv"
            return fam, 2.0 * M * N * T * Cc, 4.0 * (M * N + N * T * Cc + M * Cc)

        def f_wgrad(dy, x, dw, *, Cout, T, Cc, **k):
            Kt = dy.nb * dy.rows
            fam = "wgrad_c1" if Cc == 1 else "wgrad"
            return fam, 2.0 * Kt * Cout * T * Cc, 4.0 * (Kt * Cout + Kt * Cc + Cout * T * Cc)

        def f_gru_f(gi, w_hh, b_hh, h_out, ldh, save, B, T, H):
            return "gru_forward", 2.0 * B * T * 3 * H * H, 4.0 * B * T * (3 * H + H + (4 * H if save is not None else 0))

        def f_gru_b(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H):
            return "gru_backward", 2.0 * B * T * 3 * H * H, 4.0 * B * T * (H + H + 4 * H + 6 * H)

        def f_adam(p, g, m, v, n, *a, **k):
            return "adam", 0.0, 28.0 * n

        def f_adam_pack(table, n, smem_floats, counters, lr, *a, nbytes=0, **k):
            return "adam", 0.0, float(nbytes)       # Adam fused with the weight re-layouts: 28 B + 12 B per packed layout per parameter

        def f_dg1(dy, w, dx, *, nb, Lout, Cout, k, stride, pad, Lin):
            return "conv_dgrad_c1", 2.0 * nb * Lout * Cout * k, 4.0 * nb * (Lout * Cout + Lin)
        self._wrap("rowconv", f_rowconv)
        self._wrap("wgrad", f_wgrad)
        self._wrap("gru_forward", f_gru_f)
        self._wrap("gru_backward", f_gru_b)
        self._wrap("adam", f_adam)
        self._wrap("adam_pack", f_adam_pack)
        self._wrap("conv_dgrad_c1", f_dg1)

    def remove(self):
        for n, f in self.saved.items():
            setattr(self.ops, n, f)

    def summary(self):
        fam = {}
        for f, fl, nb, e0, e1 in self.rec:
            d = fam.setdefault(f, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += nb
        return fam


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from music2dance_b200 import config as O       # configuration + synthetic inputs (the product arm never imports oracle/)
    from music2dance_b200 import ops
    from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator
    from music2dance_b200.trainer import Phase3Trainer

    cfg = O.make_cfg(enc_type=args.enc)
    B, nc = args.batch, cfg["n_critic_steps"]
    ops.set_gemm_mode(args.gemm)
    torch.manual_seed(0)
    gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"],
                            cfg["size"], cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"],
                            cfg["n_cells"], cfg["enc_type"], cfg["activ"], dev)
    critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                                   init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
    tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=not args.no_graphs, collective=args.collective)

    # synthetic inputs of the reference shape: NSETS distinct step-input sets, pinned on the host and
    # mirrored in HBM; every rank draws its own shard (weak scaling)
    NSETS = 4
    host, devs = [], []
    for s in range(NSETS):
        bs = [O.synthetic_batch(cfg, B, 1234 + 7919 * rank + s * nc + i) for i in range(nc)]
        hs = [torch.stack([b[j] for b in bs]).pin_memory() for j in range(4)] + [bs[-1][4].pin_memory()]
        host.append(hs)
        devs.append([t.to(dev) for t in hs])
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = tr.log_c.numel() * 4 + tr.log_g.numel() * 4
    flush = torch.empty(2 * L2_BYTES // 4, dtype=torch.float32, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def step_resident(i):
        tr.load_batches(*devs[i % NSETS])
        flush.zero_()                                   # L2 flush between steps (252 MiB write)
        tr.train_step()

    def step_e2e(i):
        tr.load_batches(*host[i % NSETS])               # H2D of this step's inputs (pinned)
        flush.zero_()
        tr.train_step()
        return tr.logs()                                # D2H read of the step's scalars (synchronises)

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sync_all()
    launches = tr.launches_per_step()                   # counted C-ABI kernel launches of one step
    sync_all()

    def timed(fn):
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for i in range(args.steps):
            fn(i)
        e1.record()
        sync_all()
        clocks = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    ms_res, clocks = timed(step_resident)
    for i in range(2):
        step_e2e(i)
    ms_e2e, clocks_e2e = timed(step_e2e)
    logs = tr.logs()

    # Input pipeline on a device-resident dataset (music2dance_b200/data.py): the reference's
    # __getitem__ / collate_fn / H2D become one gather kernel per iteration driven by B sequence indices and B
    # start frames; noise and alpha are still drawn on the host (reference RNG order) and copied.
    e2e_ds = None
    if not args.no_device_dataset:
        import numpy as np
        from music2dance_b200.data import DeviceSequenceDataset
        rs = np.random.RandomState(99 + rank)
        nseq = 61                                            # size of the reference dataset (train.py:114-125)
        seqs, mus = [], []
        for _ in range(nseq):
            L = int(rs.randint(750, 3000))                   # 30 s .. 2 min at 25 fps
            seqs.append(rs.rand(L, 23, 3).astype(np.float32))
            mus.append(((rs.rand(L * 640) * 2 - 1) * 0.3).astype(np.float32))
        labels = rs.randint(0, 4, size=nseq)
        dsd = DeviceSequenceDataset(dict(sequences=seqs, musics=mus, labels=labels, dirs=[str(i) for i in range(nseq)]),
                                    dict(audio_rate=16000, video_rate=25, seq_length=4.8, feat_size=69), dev)
        count = np.unique(labels, return_counts=True)[1]
        wts = torch.as_tensor((1.0 / count)[labels], dtype=torch.double)
        T_, O_ = cfg["stick_length"], cfg["output_size"]

        def step_e2e_ds(i):
            hs = host[i % NSETS]
            for it in range(nc):
                bi = torch.multinomial(wts, B, True).tolist()          # class-balanced sampler (train.py:133-138)
                st = [dsd.positions(j)[0] for j in bi]                 # get_positions: numpy global generator
                dsd.crop(bi, st, real=tr.in_real[it].view(B, T_, 23, 3), audio=tr.in_audio[it])
            tr.in_noise.copy_(hs[2].reshape(tr.in_noise.shape), non_blocking=True)
            tr.in_alpha.copy_(hs[3].reshape(tr.in_alpha.shape), non_blocking=True)
            tr.in_noise_g.copy_(hs[4].reshape(tr.in_noise_g.shape), non_blocking=True)
            flush.zero_()
            tr.train_step()
            return tr.logs()

        for i in range(2):
            step_e2e_ds(i)
        ms_ds, _ = timed(step_e2e_ds)
        h2d_ds = nc * B * 2 * 4 + sum(host[0][j].numel() * host[0][j].element_size() for j in (2, 3, 4))
        e2e_ds = {"value": world * args.steps / (ms_ds * 1e-3), "unit": UNIT, "ms_per_step": ms_ds / args.steps,
                  "h2d_bytes_per_step": h2d_ds, "d2h_bytes_per_step": d2h,
                  "dataset": f"{nseq} synthetic sequences resident in HBM ({(dsd.poses.numel() + dsd.music.numel()) * 4 >> 20} MiB), "
                             "random crops gathered on the device (m2d_crop_batch), sampler and crop positions drawn on the "
                             "host in the reference's RNG order"}

    # flush cost (excluded from nothing: it is inside both timed regions; reported for transparency)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(10):
        flush.zero_()
    e1.record()
    torch.cuda.synchronize(dev)
    flush_ms = e0.elapsed_time(e1) / 10

    roof = None
    fams = None
    pk = peaks()
    if rank == 0 and not args.no_roofline:
        # one more train step, eager, with CUDA events around every GEMM-family launch
        kt = KernelTimer(ops, torch)
        kt.install()
        saved_flag = tr.use_graphs
        tr.use_graphs = False
        saved_ov = (tr.overlap, tr.D.par, tr.G.par)
        tr.overlap, tr.D.par, tr.G.par = False, False, False   # one stream: every launch is timed alone
        tr.load_batches(*devs[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        if world == 1:
            tr.train_step()
        else:                                            # no collectives on a single rank: kernels only
            for i in range(nc):
                tr.critic_iteration(i, update=False)
            tr.generator_update(update=False)
        e1.record()
        torch.cuda.synchronize(dev)
        eager_ms = e0.elapsed_time(e1)
        tr.use_graphs = saved_flag
        tr.overlap, tr.D.par, tr.G.par = saved_ov
        kt.remove()
        fams = kt.summary()
        top = max(fams, key=lambda f: fams[f]["ms"])
        d = fams[top]
        tensor_bound = d["flops"] / max(d["bytes"], 1.0) > 100.0
        tfp = measure_tf32_peak(torch, dev) if tensor_bound else None
        if tensor_bound:
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            peak = tfp["tf32_tflops_sustained"]          # kernels timed inside a long step: the sustained figure
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "tf32_peak_measured": tfp}
        else:
            ach = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            peak = pk["hbm_gbs"]
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None}
        # DRAM bytes per launch of the same kernel family from the committed ncu capture of THIS tree
        # (tools/gpu_traffic.sh -> tools/summarize_traffic.py); null when the capture is stale (its launch count of
        # the family differs from the one just counted) or does not cover the dominant family
        try:
            with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
                tj = json.load(f)
            key = f"{top}:b{args.batch}:{args.enc}:{args.gemm}"
            if key in tj and tj[key]["launches_per_step"] == d["launches"]:
                roof["traffic"] = tj[key]["traffic_bytes_per_launch"]
                roof["traffic_source"] = (f"profiles/{TRAFFIC_FILE}: mean dram__bytes_read+write per launch over the "
                                          f"{tj[key]['launches_per_step']} {top} launches of one train step (ncu, same tree)")
        except (OSError, ValueError, KeyError):
            pass
        roof.update({"kernel": top, "launches_per_step": d["launches"],
                     "avg_launch_us": 1e3 * d["ms"] / d["launches"],
                     "share_of_eager_step": d["ms"] / eager_ms,
                     "measured_in": "one instrumented eager train step, single stream (multi-stream overlap off), "
                                    "CUDA events around every launch of the family",
                     "peak_source": (("measured in this run: " + tfp["how"] + "; gemm mode " + args.gemm +
                                      {"tf32x3": " issues 3 tensor-core products per algorithmic product",
                                       "tf32bf16": " issues 2 TF32-equivalent tensor-core products per algorithmic product "
                                                   "(row convolutions; weight gradients 3)"}.get(args.gemm, ""))
                                     if tensor_bound else pk["source"]),
                     "algorithmic_gflop_per_launch": d["flops"] / d["launches"] / 1e9,
                     "algorithmic_bytes_per_launch": d["bytes"] / d["launches"]})
        for f in fams.values():
            f["tflops"] = f["flops"] / max(f["ms"], 1e-9) / 1e9
            f["gbs"] = f["bytes"] / max(f["ms"], 1e-9) / 1e6
            f["share_of_eager_step"] = f["ms"] / eager_ms
            for k in ("flops", "bytes"):
                f[k] = float(f[k])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t_step, sample, cores, kind, _ = reference_time(cfg, B, "cpu", 3, 1, 30.0)
        cpu = {"value": 1.0 / t_step, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    # The bar a user of the reference gets on this GPU today: the reference's stock modules through
    # PyTorch eager + cuDNN/cuBLAS on cuda:0, fp32 (TF32 off) and with TF32 allowed.  Child processes, after every timed
    # region of this run; never the product path.
    eager = None
    if rank == 0 and world == 1 and not args.sub and not args.no_eager_gpu:
        eager = {}
        for tag, extra in (("fp32", []), ("tf32", ["--ref-tf32"])):
            ln = child_line(["--impl", "reference", "--ref-device", "cuda", "--batch", str(B), "--enc", args.enc,
                             "--steps", "10", "--warmup", "3", "--cpu-budget-s", "60"] + extra, 240)
            eager[tag] = ({"value": ln["value"], "unit": UNIT, "ms_per_step": ln["ms_per_step"], "steps": ln["steps"],
                           "warmup": ln["warmup"], "kind": ln["cpu_baseline"]["kind"], "device": ln["config"]["device"],
                           "last_step_logs": ln.get("detail", {}).get("logs")} if "value" in ln else ln)
    regime = None
    if rank == 0 and world == 1 and not args.sub and not args.no_throughput_regime and args.regime_batch != B:
        ln = child_line(["--batch", str(args.regime_batch), "--enc", args.enc, "--gemm", args.gemm, "--steps", "4",
                         "--warmup", "3", "--sub", "--no-cpu-baseline", "--no-device-dataset"], 400)
        if "value" in ln:
            regime = {"batch_per_gpu": args.regime_batch, "value": ln["value"], "unit": UNIT,
                      "ms_per_step": ln["ms_per_step"], "sequences_per_s": ln["sequences_per_s"],
                      "e2e": ln["e2e"], "roofline": ln.get("roofline"), "clocks": ln.get("clocks"),
                      "kernel_families": {k: {kk: v[kk] for kk in ("launches", "ms", "tflops", "gbs", "share_of_eager_step")}
                                          for k, v in (ln.get("kernel_families") or {}).items()}}
        else:
            regime = ln

    if rank == 0:
        v = world * args.steps / (ms_res * 1e-3)
        e = world * args.steps / (ms_e2e * 1e-3)
        wk_bytes = tr.G.wk.bytes() + tr.D.wk.bytes()
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tf32": "tf32 (fp32 accumulate)",
                          "tf32bf16": "tf32 hi*hi + bf16 cross terms on tcgen05, fp32 accumulate (fp32-grade; weight "
                                      "gradients 3xTF32)",
                          "tf32x3": "tf32x3 (3xTF32 operand split on tcgen05, fp32 accumulate; fp32-grade)"}[args.gemm],
                "data": "synthetic",
                "config": {"workload": workload_name(args, cfg), "global_batch": B * world, "gemm": args.gemm,
                           "sequences_per_step": (nc) * B * world, "parallelism": f"dp{world}",
                           "collective": (None if world == 1 else
                                          ("m2d_nvl_allreduce: NVLink peer-memory two-shot all-reduce kernel inside the step "
                                           "graph (" + ("multimem.ld_reduce / multimem.st, NVLink SHARP" if tr.nvl.multicast
                                                        else "peer loads / stores, no multicast mapping") + ")")
                                          if tr.nvl is not None else
                                          "NCCL all-reduce (torch.distributed) between per-iteration graphs" +
                                          (f"; peer-memory path unavailable: {tr.nvl_error}" if getattr(tr, "nvl_error", None) else "")),
                           "cuda_graphs": not args.no_graphs,
                           "graph_structure": ("eager" if args.no_graphs else "one graph per critic iteration, NCCL "
                                               "all-reduce between graphs" if tr.per_iter else
                                               "one graph per train step"),
                           "l2": (f"L2 flushed between steps by a {2 * L2_BYTES >> 20} MiB write ({flush_ms:.3f} ms, inside the "
                                  f"timed region); inputs rotate over {NSETS} staged sets; per-iteration working set "
                                  f"(weights + Adam moments {(tr.de.fp.n_live_padded * 16) >> 20} MiB + activations "
                                  f"{wk_bytes >> 20} MiB) exceeds the 126 MiB L2")},
                "e2e": {"value": e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
                "clocks": clocks, "clocks_e2e": clocks_e2e,
                "sequences_per_s": v * nc * B,
                "value_definition": ("train steps/s summed over the ranks: every rank runs batch-B train steps on its own "
                                     "shard with all-reduced gradients (weak scaling), i.e. optimizer steps/s x n_gpus; "
                                     "sequences_per_s = value x n_critic x batch is the size-independent companion"),
                "optimizer_steps_per_s": v / world,
                "last_step_logs": {"loss_critic": logs["critic"][-1]["loss_critic"], "gp": logs["critic"][-1]["gp"],
                                   "loss_gen": logs["gen"]["loss_gen"]}}
        if e2e_ds is not None:
            line["e2e_device_dataset"] = e2e_ds
        if roof is not None:
            line["roofline"] = roof
            line["kernel_families"] = fams
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if eager is not None:
            line["torch_eager_gpu"] = eager
        if regime is not None:
            line["throughput_regime"] = regime
        print(json.dumps(line), flush=True)
    if world > 1:
        if tr.nvl is not None:
            tr.nvl.check()
        torch.cuda.synchronize(dev)
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
