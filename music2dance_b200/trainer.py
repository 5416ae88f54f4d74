"""Fused phase3 WGAN-GP train step (reference: phase3/train.py:184-243).

One train step = `n_critic_steps` critic iterations (each on its own batch: generator
forward with train-mode BatchNorm, gradient penalty, Wasserstein terms, Adam) followed by
one generator update on the last batch (Q7).  Everything between the host->device copy of
a batch and the scalar log runs in libm2d_b200 kernels on channels-last activations:

  * audio windowing is fused into the first encoder convolution (never materialised),
  * the critic's audio branch is evaluated once per iteration and shared by the
    interpolated / real / fake evaluations (Q13),
  * the gradient penalty's weight gradients come from one tangent pass + weight-gradient
    GEMMs (wgan.py) instead of autograd's double backward,
  * Adam is one flat kernel per network, followed by the weight re-layouts,
  * with world_size > 1 the flat gradient buffer is all-reduced over NCCL before Adam
    (batch data parallelism; per-replica generator BatchNorm statistics),
  * the whole step is captured in CUDA graphs (one launch per iteration).
"""
from __future__ import annotations

import os

import torch

from . import dp, ops
from .engine import AdamPack
from .ops import Mat
from .nets import ACT_ID, PRIO_CHAIN, PRIO_GEN, Workspace, make_stream
from .wgan import (critic_backward_fused, critic_forward, critic_forward_fused, gradient_penalty_pass, rows,
                   wasserstein_backward)

# one batched generator pass serves at most this many sequences (trainer.gen_groups)
GEN_GROUP_SEQUENCES = 64


def plan_gen_groups(nc, B, spec=None):
    """{first iteration of a pass: iterations it serves} for the generator forwards of the nc critic iterations of a
    train step (host logic, no device).  spec: "8", "1,7", ... (sizes in order, summing to nc); default: consecutive
    passes of at most GEN_GROUP_SEQUENCES // B iterations (at least one)."""
    if spec:
        sizes = [int(x) for x in spec.split(",")]
        if not (all(g >= 1 for g in sizes) and sum(sizes) == nc):
            raise ValueError(f"M2D_GEN_GROUPS={spec!r}: positive sizes that sum to n_critic_steps={nc} expected")
    else:
        gmax = max(1, GEN_GROUP_SEQUENCES // B)
        sizes, left = [], nc
        while left > 0:
            sizes.append(min(gmax, left))
            left -= sizes[-1]
    plan, i = {}, 0
    for g in sizes:
        plan[i] = g
        i += g
    return plan


LOG_CRITIC = ("loss_critic", "gp", "w_dist", "err_real", "err_fake")
LOG_GEN = ("loss_gen", "l1", "tv", "err_real", "err_fake")


class Phase3Trainer:
    def __init__(self, gen, critic, cfg, batch_size, use_graphs=True, process_group=None, per_iteration_graphs=False,
                 collective=None):
        """collective (world_size > 1): "nvl" = gradient all-reduce by the hand-written NVLink peer-memory kernel
        (dp.NvlAllReduce; the whole step stays ONE CUDA graph), "nccl" = torch.distributed all-reduce between
        per-iteration graphs, None = "nvl" when symmetric memory is available on every rank, else "nccl"."""
        dev = next(gen.parameters()).device
        assert dev.type == "cuda", "Phase3Trainer needs the modules on a CUDA device"
        self.dev, self.cfg, self.B = dev, cfg, batch_size
        self.T, self.O, self.A = cfg["stick_length"], cfg["output_size"], cfg["audio_length"]
        self.Nz, self.nc = cfg["noise_size"], int(cfg["n_critic_steps"])
        gen.cutting_stride, gen.pad_samples = cfg["cutting_stride"], cfg["pad_samples"]
        gen.__dict__.pop("_m2d_engine", None)
        self.gen, self.critic = gen, critic
        self.pg = process_group
        self.world = dp.world_size(process_group)
        self.nvl = self._make_nvl(collective, dev) if self.world > 1 else None
        from . import engine as _engine_mod
        with torch.cuda.device(dev):
            if self.nvl is not None:                     # gradients are written straight into node-mapped buffers
                critic.__dict__.pop("_m2d_engine", None)
                _engine_mod.GRAD_ALLOC[0] = self.nvl.alloc
            try:
                self.ge, self.de = gen._engine(), critic._engine()
            finally:
                _engine_mod.GRAD_ALLOC[0] = None
            self.G, self.D = self.ge.net, self.de.net
            if self.world > 1:
                # every replica starts from rank 0's parameters and BatchNorm buffers (identical seeds make this a
                # no-op; a resumed or re-seeded rank must not silently diverge)
                import torch.distributed as dist
                with torch.no_grad():
                    for t in [self.ge.fp.flat, self.de.fp.flat] + [b for _, b in gen.named_buffers()]:
                        dist.broadcast(t, 0, group=process_group)
            self.ge.net.pack()
            self.de.net.pack()
        self.ge.packed_version, self.de.packed_version = self.ge.fp.version(), self.de.fp.version()
        # NCCL runs replay one graph per critic iteration (the all-reduce sits between graphs); the flag selects that
        # structure on a single GPU too (tests exercise it without a second device).  With the peer-memory kernel the
        # collective is just another launch, so the multi-GPU step is ONE graph like the single-GPU one.
        self.per_iter = per_iteration_graphs or (self.world > 1 and self.nvl is None)
        B, T, O, A, Nz, nc = self.B, self.T, self.O, self.A, self.Nz, self.nc
        f = dict(dtype=torch.float32, device=dev)
        # staged inputs of one train step (device resident)
        self.in_real = torch.zeros(nc, B, T, O, **f)
        # audio staged as the first half of a [2B, A] block per iteration: the second half receives the penalty's
        # direction v1, which makes the block the stacked input operand of audio_d.l1's weight gradient (no copy)
        self.in_audio2 = torch.zeros(nc, 2 * B, A, **f)
        self.in_audio = self.in_audio2[:, :B]
        self.in_noise = torch.zeros(nc, B, T, Nz, **f)
        self.in_alpha = torch.zeros(nc, B, **f)
        self.in_noise_g = torch.zeros(B, T, Nz, **f)
        self.log_c = torch.zeros(nc, 8, **f)
        self.log_g = torch.zeros(8, **f)
        self.fake_c = torch.zeros(nc, B * T, O, **f)   # generated poses of every critic iteration
        self.fake_g = torch.zeros(B * T, O, **f)       # ... of the generator update
        nD, nG = self.de.fp.n_live_padded, self.ge.fp.n_live_padded
        self.mD, self.vD = torch.zeros(nD, **f), torch.zeros(nD, **f)
        self.mG, self.vG = torch.zeros(nG, **f), torch.zeros(nG, **f)
        # optimiser step fused with the weight re-layouts (m2d_adam_pack): one table per launch.  The critic's two
        # big late layers (audio_d.l5 / l6, 68 % of its parameters) have their own table so that their update runs on
        # a side stream next to the next iteration's first convolutions (CriticNet.pack_late_fork protocol)
        with torch.cuda.device(dev):
            self.apG = AdamPack(self.ge.fp, self.G, self.mG, self.vG)
            if self.D.can_split_pack():
                # "late" parameters: everything the next forward reaches only after audio_d.l4 — audio_d.l5 / l6 (68 % of
                # the critic) and the fusion MLP, weights and biases.  Their gradients are the tails of both gradient
                # buffers (parameter order), so each side is one contiguous range for Adam and for the all-reduce.
                fp = self.de.fp
                late = [self.D.a_layers[4], self.D.a_l6, self.D.fc1, self.D.fc2]
                base = fp.flat.data_ptr()
                off = lambda t: (t.data_ptr() - base) // 4
                self._late_plain = off(self.D.a_layers[4].b)
                tail = {off(t) for c in late for t in (c.b,)} | {off(self.D.fc1.w), off(self.D.fc2.w)}
                plain_offs = [off(p) for p in fp.params.values() if off(p) < fp.n_plain_padded]
                assert {o for o in plain_offs if o >= self._late_plain} == tail and self._late_plain % 4 == 0
                self.apD = AdamPack(fp, self.D, self.mD, self.vD, exclude=late, flat_range=(0, self._late_plain))
                self.apD_late = AdamPack(fp, self.D, self.mD, self.vD, only=late,
                                         flat_range=(self._late_plain, fp.n_plain_padded))
                self._late_off = (late[0].gwp.data_ptr() - fp.gpk.data_ptr()) // 4
                assert late[1].gwp.data_ptr() == late[0].gwp.data_ptr() + 4 * ((late[0].gwp.numel() + 3) // 4 * 4)
                assert self._late_off % 4 == 0
            else:
                self.apD = AdamPack(self.de.fp, self.D, self.mD, self.vD)
                self.apD_late = None
        self.gp_buf = torch.zeros(1, **f)
        self.k0, self.k1 = torch.zeros(B, **f), torch.zeros(B, **f)
        self.use_graphs = use_graphs
        self.graphs = None
        # The generator's weights do not change during the n_critic critic iterations, so its forwards
        # (one per iteration + the one of the generator update, in the reference's order: BatchNorm
        # running statistics advance sequentially) run on a side stream and overlap the critic work.
        self.overlap = os.environ.get("M2D_OVERLAP", "1") != "0"
        self.s_gen = make_stream(dev, PRIO_GEN)
        self.s_main = make_stream(dev, PRIO_CHAIN)          # capture stream of the step graph(s)
        # ... and because the weights are constant, several of these forwards can be ONE batched pass with per-batch
        # BatchNorm statistics (GeneratorNet.forward groups=): gen_groups maps the first iteration of a pass to the
        # number of iterations it serves.  Small batches only (the kernels of a batch-7 forward are latency-bound; from
        # 64 sequences on they fill the machine anyway).  M2D_GEN_GROUPS="1,7" style lists override the default.
        # Measured on B200, batch 7, n_critic 8 (train steps/s): one pass per iteration 59.1, "1,7" 61.2, "1,3,4" 61.3,
        # "8" (default) 61.7; with "8", M2D_GRU_BG 2 / 4 / 8: 60.8 / 61.7 / 60.7.
        self.gen_groups = self._plan_gen_groups(os.environ.get("M2D_GEN_GROUPS"))
        self.gen_wk = {g: Workspace(dev, scratch_floats=1 << 25) for g in set(self.gen_groups.values()) if g > 1}
        self.gen_audio = torch.zeros(nc * B, A, **f) if self.gen_wk else None
        # measured on B200 at batch 7: 0 -> 57.0, 4 -> 55.8, 8 -> 50.3 train steps/s: under contention a generator
        # forward takes about as long as a critic iteration, so its latency matters as much as its SM footprint
        self.gru_bg = int(os.environ.get("M2D_GRU_BG", "0"))
        self.split_pack = True          # critic re-layout of the late layers on a side stream (CriticNet.pack)
        # one backward sweep per critic iteration (wgan.critic_backward_fused); M2D_FUSED_BWD=0: two chains
        self.fused_backward = os.environ.get("M2D_FUSED_BWD", "1") != "0"

    # ------------------------------------------------------------------ pieces
    def _make_nvl(self, collective, dev):
        """dp.NvlAllReduce if requested / available on EVERY rank (agreed on through one NCCL all-reduce), else None."""
        import torch.distributed as dist
        if collective == "nccl":
            return None
        nvl, err = None, None
        try:
            with torch.cuda.device(dev):
                nvl = dp.NvlAllReduce(self.pg, dev)
                probe = nvl.alloc(1024)                      # exercises allocation + rendezvous on every rank
                probe.fill_(1.0)
                torch.cuda.synchronize(dev)
                dist.barrier(group=self.pg)
                nvl.all_reduce_sum_(probe, slot=0, blocks=2)
                torch.cuda.synchronize(dev)
                nvl.check()
                if abs(float(probe.sum().item()) - 1024.0 * self.world) > 1e-3:
                    raise RuntimeError(f"peer-memory all-reduce probe returned {float(probe.sum().item())}")
        except Exception as e:                                # noqa: BLE001
            nvl, err = None, e
        ok = torch.tensor([1 if nvl is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)
        if int(ok.item()) == 0:
            if collective == "nvl":
                raise RuntimeError(f"collective='nvl' requested but the peer-memory path is unavailable: {err!r}")
            self.nvl_error = repr(err)
            return None
        return nvl

    def _nvl_reduce(self, eng, part):
        """Peer-memory all-reduce of one network's gradient buffers.  part: "all", or for the critic with split tables
        "early" / "late" (the ranges of the two Adam tables; each is ONE launch over both gradient buffers)."""
        plain, *rest = eng.fp.grad_buffers()
        base = 0 if eng is self.de else 4
        if not rest:
            self.nvl.all_reduce_sum_(plain, slot=base, blocks=4)
        elif part == "all":
            self.nvl.all_reduce_sum2_(rest[0], plain, slot=base)
        elif part == "early":      # everything the early Adam table reads, in ONE launch
            self.nvl.all_reduce_sum2_(rest[0][:self._late_off], plain[:self._late_plain], slot=base + 1)
        else:                       # "late": audio_d.l5 / l6 + fusion MLP
            self.nvl.all_reduce_sum2_(rest[0][self._late_off:], plain[self._late_plain:], slot=base + 2)

    def _all_reduce(self, eng):
        """NCCL path: sum of the gradient buffers an optimiser step reads (engine.FlatParams.grad_buffers) over the
        ranks, issued between the per-iteration graphs."""
        if self.world > 1 and self.nvl is None:
            for t in eng.fp.grad_buffers():
                dp.all_reduce_sum_(t, self.pg)

    def _adam(self, eng, lr, late_fork=None, early_reduced=False):
        """Adam + re-layout of one network (gradients are read where the kernels left them: tap-major for the
        convolution weights).  late_fork: run the critic's late table on the re-layout side stream (True), inline
        (False), or decide from the step structure (None)."""
        gs = 1.0 / self.world
        nvl = self.nvl is not None
        if eng is self.ge:
            if nvl:
                self._nvl_reduce(eng, "all")
            self.apG.step(lr, gs)
            ops.mark("adam_pack")
            return
        if self.apD_late is not None:
            def late():
                if nvl:
                    self._nvl_reduce(eng, "late")
                self.apD_late.step(lr, gs)
            fork = (not self.per_iter and self.overlap and self.split_pack) if late_fork is None else late_fork
            if fork:
                self.D.late_fork(late)
            else:
                late()
        if nvl and not early_reduced:
            self._nvl_reduce(eng, "all" if self.apD_late is None else "early")
        self.apD.step(lr, gs)
        ops.mark("adam_pack")

    def _gru_side(self, on):
        """The generator forwards run on a side stream next to the critic iterations; M2D_GRU_BG = 1..8 lets their GRU
        recurrences serve that many sequences per thread-block cluster (fewer SMs busy, longer steps).  Default 0 = the
        latency-optimal plan (see __init__)."""
        if self.overlap and self.gru_bg:
            ops.set_gru_forward_batch_group(self.gru_bg if on else 0)

    def _plan_gen_groups(self, spec):
        return plan_gen_groups(self.nc, self.B, spec)

    def _gen_forward(self, i):
        """Generator forward(s) of critic iteration i (train-mode BatchNorm, nothing kept for a backward): the pass
        that STARTS at iteration i serves gen_groups[i] iterations; for the others the poses are already there."""
        g = self.gen_groups.get(i)
        if g is None:
            return
        B, T = self.B, self.T
        self._gru_side(True)
        if g == 1:
            self.G.forward(self.in_audio[i], self.in_noise[i], B, T, train=True,
                           out=Mat.of(self.fake_c[i], 1, B * T, self.O))
        else:
            # audio of the g iterations, contiguous (it is staged as the first half of [2B, A] blocks): one 2-D copy
            dst = self.gen_audio[i * B:(i + g) * B]
            ops.copy2d(Mat(self.in_audio2[i:i + g], 1, g, B * self.A, 2 * B * self.A), Mat(dst, 1, g, B * self.A))
            self.G.forward(dst, self.in_noise[i:i + g], g * B, T, train=True,
                           out=Mat.of(self.fake_c[i:i + g], 1, g * B * T, self.O), wk=self.gen_wk[g], groups=g)
        self._gru_side(False)

    def _gen_forward_update(self):
        """Generator forward of the generator update (its activations feed G.backward)."""
        self._gru_side(True)
        self.G.forward(self.in_audio[self.nc - 1], self.in_noise_g, self.B, self.T, train=True,
                       out=Mat.of(self.fake_g, 1, self.B * self.T, self.O))
        self._gru_side(False)

    def critic_iteration(self, i, update=True, gen_inline=True, fake_ready=None):
        """train.py:187-216 on staged batch i.  fake_ready: event of the (side-stream) generator forward of this
        iteration; the fused path waits for it only after the audio branch — which needs neither the generated poses
        nor the interpolates — has been forked."""
        B, T, O, D, G = self.B, self.T, self.O, self.D, self.G
        real, audio = self.in_real[i], self.in_audio[i]
        if gen_inline:
            self._gen_forward(i)
        ops.mark(f"it{i}:start")
        fake_c = self.fake_c[i]
        wk = D.wk
        wk.acc_reset()
        n3 = 3 * B
        X3 = wk.mat("c:X3", n3, T, O)
        per = T * O
        def stack():
            if fake_ready is not None:
                torch.cuda.current_stream(self.dev).wait_event(fake_ready)
            ops.interp_stack3(real, fake_c, self.in_alpha[i], X3, B, per)      # [interpolates; real; fake]
        gamma = float(self.cfg["gamma"])
        fused = self.fused_backward and D.act == ACT_ID
        if not fused:
            stack()
        if fused:
            # one backward sweep for the Wasserstein terms and the penalty (wgan.critic_backward_fused)
            fw = critic_forward_fused(D, X3, None if D.ablated else audio, B, "c", before_pose=stack)
            d = fw["d"]
            aud2 = None if D.ablated else Mat(self.in_audio2[i], 2 * B, self.A, 1)
            hook = None
            if update and self.nvl is not None and self.apD_late is not None:
                hook = lambda: self._nvl_reduce(self.de, "early")
            early = critic_backward_fused(D, fw, B, audio, gamma, self.gp_buf, self.k0, self.k1, aud2=aud2,
                                          early_reduce=hook)
            ops.wgan_scalars(None, self.gp_buf, B, 1, 1, gamma, 0.0, 0, self.log_c[i],
                             d_real=rows(d, B, 2 * B), d_fake=rows(d, 2 * B, n3))
            if update:
                self._adam(self.de, self.cfg["lr_critic"], early_reduced=bool(early))
            elif not self.per_iter:
                D.unpack_grads()              # caller wants the gradients in the parameter layout (tests, tools)
            return
        fw = critic_forward(D, X3, None if D.ablated else audio, n3, B, "c", groups=3)
        sums = wk.acc_slot(4)
        d = fw["d"]
        ops.sum_(rows(d, B, 2 * B), B, sums[0:1])
        ops.sum_(rows(d, 2 * B, n3), B, sums[1:2])
        # two independent backward chains over the same forward state: the Wasserstein terms (second
        # stream pair; they OVERWRITE the gradient buffers) and the gradient penalty (backward-data,
        # tangent pass; its weight gradients ACCUMULATE once the first chain is done)
        with D.fork_w():
            wasserstein_backward(D, fw, B, 2 * B, (-1.0, 1.0), B, "c:w", beta=0.0)
        gradient_penalty_pass(D, fw, B, "c:gp", gamma, 1.0, self.gp_buf, self.k0, self.k1, before_wgrads=D.join_w)
        ops.wgan_scalars(sums, self.gp_buf, B, 1, 1, gamma, 0.0, 0, self.log_c[i])
        if update:
            self._adam(self.de, self.cfg["lr_critic"])
        elif not self.per_iter:
            D.unpack_grads()

    def generator_update(self, update=True, gen_inline=True):
        """train.py:222-237 on the last staged batch."""
        B, T, O, D, G = self.B, self.T, self.O, self.D, self.G
        i = self.nc - 1
        real, audio = self.in_real[i], self.in_audio[i]
        if gen_inline:
            self._gen_forward_update()
        ops.mark("gu:start")
        wk = D.wk
        wk.acc_reset()
        n2 = 2 * B
        X2 = wk.mat("g:X2", n2, T, O)
        per = T * O
        ops.axpby(real, None, rows(X2, 0, B), B * per, 1.0, 0.0)
        ops.axpby(self.fake_g, None, rows(X2, B, n2), B * per, 1.0, 0.0)
        fw = critic_forward(D, X2, None if D.ablated else audio, n2, B, "g", groups=2)
        sums = wk.acc_slot(4)
        d = fw["d"]
        ops.sum_(rows(d, 0, B), B, sums[0:1])
        ops.sum_(rows(d, B, n2), B, sums[1:2])
        dfake = wk.mat("g:dfake", B, T, O)
        wasserstein_backward(D, fw, B, B, (-1.0,), B, "g:w", beta=0.0, dX_rows=0, dX=dfake, param_grads=False)
        beta, eta = float(self.cfg["beta"]), float(self.cfg["eta"])
        ops.pose_losses(real, self.fake_g, dfake, B, T, O, beta, eta, True, sums[2:4])
        ops.wgan_scalars(sums, None, B, B * T * O, B * (T - 1) * O, beta, eta, 1, self.log_g)
        ops.mark("gu:critic_done")
        G.backward(dfake.flat_rows())
        ops.mark("gu:gen_bwd")
        if update:
            self._adam(self.ge, self.cfg["lr_gen"])
        elif not self.per_iter:
            G.unpack_grads()

    # ------------------------------------------------------------------ step
    def _state(self):
        ts = [self.ge.fp.flat, self.de.fp.flat, self.mD, self.vD, self.mG, self.vG, self.apG.counters, self.apD.counters]
        if self.apD_late is not None:
            ts.append(self.apD_late.counters)
        ts += [b for _, b in self.gen.named_buffers()]
        return ts

    def _capture(self):
        """Warm up once eagerly (allocates every workspace buffer), then capture the step
        in CUDA graphs; model / optimiser state is restored afterwards."""
        saved = [t.clone() for t in self._state()]
        if self.D.can_split_pack():
            self.D._split_tabs()                 # device tables of the split re-layout: built outside capture
        torch.cuda.synchronize(self.dev)
        self._run_eager()
        torch.cuda.synchronize(self.dev)
        graphs = []
        if not self.per_iter:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.s_main):
                self._run_eager()
            graphs.append(("all", g))
        else:
            # NCCL all-reduce stays outside the graphs: [grads] -> all-reduce -> [adam + repack].
            # Generator forwards are software-pipelined: graph i runs the forward the NEXT iteration needs
            # (or the generator update's) on the side stream next to critic iteration i.
            def cap(fn):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.s_main):
                    fn()
                return g

            # critic re-layout split across graphs: the Adam graph refreshes everything but audio_d.l5 / l6, the NEXT
            # graph that evaluates the critic forks their re-layout at its start (joined before l5 by audio_fwd)
            split = self.overlap and self.split_pack and self.D.can_split_pack()

            def critic_and_next_forward(i):
                if not self.overlap:
                    self.critic_iteration(i, update=False)
                    return
                main = torch.cuda.current_stream()
                if split and i > 0:
                    late_fork()
                self.s_gen.wait_stream(main)
                with torch.cuda.stream(self.s_gen):
                    if i + 1 < self.nc:
                        self._gen_forward(i + 1)         # no-op unless a (batched) pass starts at iteration i + 1
                    else:
                        self._gen_forward_update()
                self.critic_iteration(i, update=False, gen_inline=False)
                main.wait_stream(self.s_gen)

            # the late critic table (audio_d.l5 / l6) is not part of the Adam graph: the NEXT graph that evaluates the
            # critic forks it at its start, so it overlaps that graph's first convolutions
            def adam_d():
                self.apD.step(self.cfg["lr_critic"], 1.0 / self.world)
                if not split and self.apD_late is not None:
                    self.apD_late.step(self.cfg["lr_critic"], 1.0 / self.world)

            def adam_g():
                self.apG.step(self.cfg["lr_gen"], 1.0 / self.world)

            def late_fork():
                self.D.late_fork(lambda: self.apD_late.step(self.cfg["lr_critic"], 1.0 / self.world))

            if self.overlap:
                graphs.append(("p", cap(lambda: self._gen_forward(0))))
            for i in range(self.nc):
                graphs.append(("c", cap(lambda i=i: critic_and_next_forward(i))))
                graphs.append(("cu", cap(adam_d)))
            def gen_update():
                if split:
                    late_fork()                              # audio_d.l5 / l6 update of the last critic Adam step
                self.generator_update(update=False, gen_inline=not self.overlap)

            graphs.append(("g", cap(gen_update)))
            graphs.append(("gu", cap(adam_g)))
        torch.cuda.synchronize(self.dev)
        with torch.no_grad():
            for t, s in zip(self._state(), saved):
                t.copy_(s)
        self.ge.net.pack()
        self.de.net.pack()
        torch.cuda.synchronize(self.dev)
        self.graphs = graphs

    def _run_eager(self):
        if not self.overlap or self.per_iter:
            for i in range(self.nc):
                self.critic_iteration(i)
            self.generator_update()
            return
        main = torch.cuda.current_stream()
        self.s_gen.wait_stream(main)                  # staged inputs, updated generator weights
        evs = []
        with torch.cuda.stream(self.s_gen):
            for i in range(self.nc):
                self._gen_forward(i)
                ev = torch.cuda.Event()
                ev.record(self.s_gen)
                evs.append(ev)
            self._gen_forward_update()
            ev_g = torch.cuda.Event()
            ev_g.record(self.s_gen)
        for i in range(self.nc):
            self.critic_iteration(i, gen_inline=False, fake_ready=evs[i])
        main.wait_event(ev_g)
        self.generator_update(gen_inline=False)

    def load_batches(self, real, audio, noise, alpha, noise_g, non_blocking=True):
        """Stage one train step's inputs (host or device tensors):
        real (nc,B,T,23,3)|(nc,B,T,69), audio (nc,B,A), noise (nc,B,T,Nz), alpha (nc,B[,1]),
        noise_g (B,T,Nz)."""
        self.in_real.copy_(real.reshape(self.in_real.shape), non_blocking=non_blocking)
        self.in_audio.copy_(audio.reshape(self.in_audio.shape), non_blocking=non_blocking)
        self.in_noise.copy_(noise.reshape(self.in_noise.shape), non_blocking=non_blocking)
        self.in_alpha.copy_(alpha.reshape(self.in_alpha.shape), non_blocking=non_blocking)
        self.in_noise_g.copy_(noise_g.reshape(self.in_noise_g.shape), non_blocking=non_blocking)

    def _ensure_packed(self):
        """Parameters changed behind the trainer's back (load_state_dict to resume, an external optimiser step):
        refresh the packed weight copies the kernels read before the next step (host-side version check)."""
        for eng in (self.ge, self.de):
            v = eng.fp.version()
            if v != eng.packed_version:
                eng.net.pack()
                eng.packed_version = v

    def optimizer_state(self):
        """Adam state for checkpoint / resume (moments laid out like the flat parameter buffers)."""
        return {"mD": self.mD.clone(), "vD": self.vD.clone(), "mG": self.mG.clone(), "vG": self.vG.clone(),
                "stepD": int(self.apD.counters[0]), "stepG": int(self.apG.counters[0])}

    def load_optimizer_state(self, st):
        with torch.no_grad():
            for k in ("mD", "vD", "mG", "vG"):
                getattr(self, k).copy_(st[k])
            self.apD.counters[0] = int(st["stepD"])
            if self.apD_late is not None:
                self.apD_late.counters[0] = int(st["stepD"])
            self.apG.counters[0] = int(st["stepG"])

    def train_step(self):
        """Run one train step on the staged inputs (asynchronous; read `logs()` to sync)."""
        with torch.cuda.device(self.dev):
            self._ensure_packed()
            if not self.use_graphs:
                self._run_eager()
                return
            if self.graphs is None:
                self._capture()
            for kind, g in self.graphs:
                g.replay()
                if kind == "c":
                    self._all_reduce(self.de)
                elif kind == "g":
                    self._all_reduce(self.ge)

    def validate(self, real, audio, noise):
        """Validation pass of phase3/train.py:245-261: eval-mode generator (BatchNorm running statistics) on one
        batch, mean |real - fake|.  real (Bv,T,23,3)|(Bv,T,69), audio (Bv,A), noise (Bv,T,Nz) (the reference draws
        it with torch.randn on the CPU generator; pass that tensor for parity).  Returns (l1 0-d tensor, fake
        (Bv*T, 69)).  Uses its own buffers, never touches optimiser state or running statistics."""
        Bv = real.shape[0]
        f = dict(dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            r = real.to(**f).reshape(Bv, self.T, self.O).contiguous()
            a = audio.to(**f).reshape(Bv, self.A).contiguous()
            z = noise.to(**f).reshape(Bv, self.T, self.Nz).contiguous()
            fake = torch.empty(Bv * self.T, self.O, **f)
            self.ge.ensure_packed()
            self.G.forward(a, z, Bv, self.T, train=False, out=Mat.of(fake, 1, Bv * self.T, self.O))
            acc = torch.zeros(2, dtype=torch.float64, device=self.dev)
            ops.pose_losses(r, fake, None, Bv, self.T, self.O, 0.0, 0.0, False, acc)
            return (acc[0] / (Bv * self.T * self.O)).float(), fake

    def logs(self):
        """Scalars of the last train step (device->host read; synchronises)."""
        c = self.log_c.cpu()
        g = self.log_g.cpu()
        out = {"critic": [dict(zip(LOG_CRITIC, c[i, :5].tolist())) for i in range(self.nc)],
               "gen": dict(zip(LOG_GEN, g[:5].tolist()))}
        return out

    def launches_per_step(self):
        """Number of libm2d kernel launches one train step issues (counted, not estimated)."""
        before = ops.LAUNCHES[0]
        saved = [t.clone() for t in self._state()]
        with torch.cuda.device(self.dev):
            self._run_eager()
            n = ops.LAUNCHES[0] - before
            torch.cuda.synchronize(self.dev)
            with torch.no_grad():
                for t, s in zip(self._state(), saved):
                    t.copy_(s)
            self.ge.net.pack()
            self.de.net.pack()
        return n
