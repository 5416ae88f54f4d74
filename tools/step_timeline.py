#!/usr/bin/env python
"""In-graph timeline of one train step: the phase marks (`ops.mark`, a one-thread kernel that stores %globaltimer)
are captured together with the step's CUDA graph(s), so the timestamps are those of the TIMED configuration
(multi-stream overlap, graph replay), not of a serialising profiler.
    python tools/step_timeline.py [batch] [out.json]
Prints, per critic iteration (median over the iterations of the last replay), the time of every mark relative to
the iteration's start, and the same for the generator update and one generator forward.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from music2dance_b200 import config as O, ops                                               # noqa: E402
from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator       # noqa: E402
from music2dance_b200.trainer import Phase3Trainer                                          # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 7
out = sys.argv[2] if len(sys.argv) > 2 else None
enc = os.environ.get("M2D_ENC", "default")
cfg = O.make_cfg(enc_type=enc)
nc = cfg["n_critic_steps"]
torch.manual_seed(0)
gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"],
                        cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], cfg["enc_type"],
                        cfg["activ"], dev)
critic = SequenceDiscriminator(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                               init_ker=cfg["init_kernel"], activ=cfg["activ"], device=dev)
tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=True)
bs = [O.synthetic_batch(cfg, B, 1234 + i) for i in range(nc)]
tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])

trace = ops.Trace(dev)


class _Cap:
    """marks are recorded only during graph capture (the eager warm-up inside _capture would double them)"""
    def __init__(self, t):
        self.t = t

    def mark(self, name):
        if torch.cuda.is_current_stream_capturing():
            self.t.mark(name)


ops.TRACE[0] = _Cap(trace)
with torch.cuda.device(dev):
    tr.train_step()                 # captures (with marks) and replays
    for _ in range(5):
        tr.train_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.train_step()
    e1.record()
    torch.cuda.synchronize()
ops.TRACE[0] = None
rec = trace.read()
t0 = min(v for _, _, v in rec)
rows = [(n, s, (v - t0) / 1e3) for n, s, v in rec]
print(f"batch {B}, encoder {enc}: step {e0.elapsed_time(e1):.3f} ms, {len(rows)} marks")

# split into segments at the it{i}:start / gu:start marks of the main chain; generator forwards at g:start
segs, cur = [], None
gsegs, gcur = [], None
for n, s, t in rows:
    if n.startswith("g:") or n.startswith("gb:"):
        if n == "g:start":
            gcur = [(n, t)]
            gsegs.append(gcur)
        elif gcur is not None:
            gcur.append((n, t))
        if not n.startswith("gb:"):
            continue
    if n.endswith(":start") and (n.startswith("it") or n.startswith("gu")):
        cur = [(n, t)]
        segs.append(cur)
    elif cur is not None:
        cur.append((n, t))


def table(title, seglist):
    if not seglist:
        return
    names = [n for n, _ in seglist[0]]
    print(f"\n{title}  ({len(seglist)} instances; us after the first mark: median [min..max])")
    for j, n in enumerate(names):
        vals = sorted(sg[j][1] - sg[0][1] for sg in seglist if len(sg) > j)
        print(f"  {n:28s} {vals[len(vals) // 2]:9.1f}  [{vals[0]:9.1f} .. {vals[-1]:9.1f}]")


its = [sg for sg in segs if sg[0][0].startswith("it")]
# normalise names (strip iteration index)
its_n = [[(n.split(":", 1)[1] if n.startswith("it") else n, t) for n, t in sg] for sg in its]
table("critic iteration (main chain + side streams)", its_n[1:] if len(its_n) > 2 else its_n)
table("generator update", [sg for sg in segs if sg[0][0].startswith("gu")])
table("generator forward", gsegs)
starts = [sg[0][1] for sg in its]
if len(starts) > 1:
    d = [b - a for a, b in zip(starts, starts[1:])]
    print("\niteration start-to-start (us):", " ".join(f"{x:.0f}" for x in d))
if out:
    with open(out, "w") as f:
        json.dump({"batch": B, "enc": enc, "step_ms": e0.elapsed_time(e1), "marks": rows}, f)
