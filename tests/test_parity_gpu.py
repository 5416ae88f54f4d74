"""GPU parity proper: the CUDA path, called through the drop-in API (archis.default /
losses / utils — the reference-facing boundary over the C ABI), against
  (1) golden fixtures produced by the UNMODIFIED reference modules, and
  (2) the CPU oracle evaluated in-test on the same seeded inputs.
Tolerances are written next to each check (tests/parity.py)."""
import copy

import pytest
import torch

from oracle import phase3_oracle as O
from tests.parity import (ALPHA_SEED, B_GOLD, TOL_CHAINED, TOL_DRIFT_BUF, TOL_DRIFT_LR, TOL_FP32, TOL_GRAD_BIAS, TOL_GEN_GRAD_E2E, TOL_GRAD, TOL_KINK_L2, TOL_KINK_MAX,
                          TOL_NORTH_STAR, VARIANTS,
                          digest_check, load_golden, scalar_check)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(cfg, state="init"):
    from music2dance_b200.archis.default import (AblatedSequenceDiscriminator, SequenceDiscriminator,
                                                 SequenceGenerator)
    torch.manual_seed(0)                                         # phase3/train.py:35
    gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"],
                            cfg["size"], cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"],
                            cfg["n_cells"], cfg["enc_type"], cfg["activ"], DEV)
    cls = AblatedSequenceDiscriminator if cfg["ablated"] else SequenceDiscriminator
    critic = cls(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                 init_ker=cfg["init_kernel"], activ=cfg["activ"], device=DEV)
    if state == "perturbed":
        for m in (gen, critic):
            sd = {k: v.cpu() for k, v in m.state_dict().items()}
            O.perturb_params(sd)
            m.load_state_dict(sd, strict=True)
    return gen, critic


def oracle_params(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


@pytest.mark.parametrize("state", ["init", "perturbed"])
@pytest.mark.parametrize("variant", ["default", "wavegan", "unet", "ablated", "tanh"])
def test_dropin_step_vs_reference_fixtures(variant, state):
    """phase3/train.py:187-237 written with the drop-in API, compared with what the
    reference modules produced for the same seeds (fixtures)."""
    from music2dance_b200.losses import gradient_penalty, tv_loss
    from music2dance_b200.utils import slice_audio_batch
    cfg = O.make_cfg(**VARIANTS[variant])
    gold = load_golden(variant)
    gen, critic = build(cfg, state)
    for m in (gen, critic):                                      # a15 + Appendix A: same keys, same init
        for k, v in m.state_dict().items():
            digest_check(v, gold, f"{state}/init/{k}", 1e-7, f"init {k}")
    B, T, Oo = B_GOLD, cfg["stick_length"], cfg["output_size"]
    real_bt, audio, noise, _, noise_g = O.synthetic_batch(cfg, B, int(gold[f"{state}/data_seed"]))
    gen.train()
    critic.zero_grad()
    slices = slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
    assert torch.equal(slices.cpu(), O.slice_audio_batch(audio, cfg["audio_feat_samples"],
                                                         cfg["cutting_stride"], cfg["pad_samples"]))
    real = real_bt.to(DEV).view(B, T, Oo).permute(0, 2, 1).contiguous()
    aud = audio.to(DEV).unsqueeze(1)
    fake = gen(slices, [T] * B, noise=noise.to(DEV)).view(B, T, Oo).permute(0, 2, 1).contiguous()
    ref_fake = torch.from_numpy(gold[f"{state}/critic/fake"])
    assert float((fake.detach().cpu() - ref_fake).abs().max()) < TOL_FP32 * float(ref_fake.abs().max())
    torch.manual_seed(ALPHA_SEED)
    if cfg["ablated"]:
        gp = gradient_penalty(critic, B, real, fake, is_seq=True, lp=False, device=DEV)
        err_real, err_fake = critic(real).mean(), critic(fake.detach()).mean()
    else:
        gp = gradient_penalty(critic, B, real, fake, aud, is_seq=True, lp=False, device=DEV)
        err_real, err_fake = critic(real, aud).mean(), critic(fake.detach(), aud).mean()
    err = err_fake - err_real + cfg["gamma"] * gp
    err.backward()
    # north-star tolerance: 1e-3 relative on losses and gradient-penalty terms; fp32 path held tighter
    scalar_check(gp, gold[f"{state}/critic/gp"], TOL_FP32, "gp")
    scalar_check(err, gold[f"{state}/critic/loss_critic"], TOL_FP32, "loss_critic")
    scalar_check(err_fake - err_real, gold[f"{state}/critic/w_dist"], TOL_FP32, "w_dist")
    for k, p in critic.named_parameters():
        assert p.grad is not None or k == "fc2.bias", k
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        digest_check(g, gold, f"{state}/critic/grad/{k}", TOL_GRAD_BIAS if k.endswith(".bias") else TOL_GRAD,
                     f"critic grad {k}", abs_floor=1e-4, kinks=True)
    for k, v in gen.state_dict().items():                       # Q2: running stats advanced by the forward
        if "running" in k or "num_batches" in k:
            digest_check(v, gold, f"{state}/critic/genbuf/{k}", TOL_FP32, f"bn buffer {k}")
    # generator update (train.py:222-237); `fake` stays a permuted view (Q8)
    gen.zero_grad()
    fake = gen(slices, [T] * B, noise=noise_g.to(DEV)).view(B, T, Oo).permute(0, 2, 1)
    l1 = torch.nn.L1Loss(reduction="mean")(real, fake)
    if cfg["ablated"]:
        err_real, err_fake = critic(real).mean(), critic(fake).mean()
    else:
        err_real, err_fake = critic(real, aud).mean(), critic(fake, aud).mean()
    tv = tv_loss(fake)
    err_g = err_real - err_fake + cfg["beta"] * l1 + cfg["eta"] * tv
    err_g.backward()
    scalar_check(err_g, gold[f"{state}/gen/loss_gen"], TOL_FP32, "loss_gen")
    scalar_check(l1, gold[f"{state}/gen/l1"], TOL_FP32, "l1")
    scalar_check(tv, gold[f"{state}/gen/tv"], TOL_FP32, "tv")
    skip = set(O.pre_bn_bias_names(dict(gen.state_dict())))
    for k, p in gen.named_parameters():
        if f"{state}/gen/nograd/{k}" in gold.files:              # Q1: dead branch never gets a gradient
            assert p.grad is None, k
        elif k not in skip:
            digest_check(p.grad, gold, f"{state}/gen/grad/{k}", TOL_GEN_GRAD_E2E, f"gen grad {k}", abs_floor=1e-4, kinks=True)


@pytest.mark.parametrize("variant", ["default", "wavegan", "unet"])
def test_generator_backward_smooth(variant):
    """Generator forward/backward against oracle autograd with a SMOOTH upstream gradient
    (no L1 sign, no critic ReLU kinks): every parameter gradient within TOL_GRAD."""
    cfg = O.make_cfg(**VARIANTS[variant])
    gen, _ = build(cfg, "perturbed")
    B, T = 3, 120
    g = torch.Generator().manual_seed(21)
    audio = (torch.rand(B, cfg["audio_length"], generator=g) - 0.5) * 0.6
    noise = torch.randn(B, T, cfg["noise_size"], generator=g)
    up = torch.randn(B * T, cfg["output_size"], generator=g)
    sl = O.slice_audio_batch(audio, cfg["audio_feat_samples"], cfg["cutting_stride"], cfg["pad_samples"])
    G0 = oracle_params(gen)
    names = O.trainable_names(G0)
    ref = {}
    for dt in (torch.float32, torch.float64):                   # fp64 = ground truth, fp32 = the reference's arithmetic
        P = O._leaf({k: (v.to(dt).clone() if v.is_floating_point() else v.clone()) for k, v in G0.items()})
        o = O.generator_forward(P, cfg, sl.to(dt), noise.to(dt), train=True)
        gl = torch.autograd.grad((o * up.to(dt)).sum(), [P[k] for k in names], allow_unused=True)
        ref[dt] = (o.detach(), dict(zip(names, gl)))
    out_ref, g64 = ref[torch.float64]
    g32 = ref[torch.float32][1]
    from music2dance_b200.utils import slice_audio_batch
    gen.train()
    out = gen(slice_audio_batch(audio.to(DEV), cfg["audio_feat_samples"], cfg["cutting_stride"],
                                cfg["pad_samples"]), [T] * B, noise=noise.to(DEV))
    err = float((out.detach().cpu().double() - out_ref).abs().max() / out_ref.abs().max())
    assert err < 2e-5, f"generator output rel err vs fp64 {err:.3e}"
    (out * up.to(DEV)).sum().backward()
    skip = set(O.pre_bn_bias_names(G0))
    got = dict(gen.named_parameters())
    worst = 0.0
    for k in names:
        gr = g64[k]
        if gr is None:
            assert got[k].grad is None, k
            continue
        if k in skip:
            continue
        d = got[k].grad.cpu().double() - gr
        e_l2 = float(d.norm() / gr.norm().clamp_min(1e-30))
        e_mx = float(d.abs().max() / gr.abs().max().clamp_min(1e-4))
        e_ref = float((g32[k].double() - gr).norm() / gr.norm().clamp_min(1e-30))
        worst = max(worst, e_l2)
        # kink-flip bound (tests/parity.py), or three times the fp32 reference's own error
        assert e_l2 < max(TOL_KINK_L2, 3 * e_ref), f"{variant} {k}: l2 {e_l2:.3e} (fp32 reference {e_ref:.3e})"
        assert e_mx < max(TOL_KINK_MAX, 3 * e_ref), f"{variant} {k}: max {e_mx:.3e}"
    print(f"[{variant}] worst generator-gradient l2 error vs fp64 oracle: {worst:.2e}")
    # eval mode (running statistics) — train.py:245-261
    gen.eval()
    with torch.no_grad():
        out_e = gen(slice_audio_batch(audio.to(DEV), cfg["audio_feat_samples"], cfg["cutting_stride"],
                                      cfg["pad_samples"]), [T] * B, noise=noise.to(DEV))
    Pe = oracle_params(gen)
    ref_e = O.generator_forward(Pe, cfg, sl, noise, train=False)
    err = float((out_e.cpu() - ref_e).abs().max() / ref_e.abs().max())
    assert err < TOL_FP32, f"eval-mode generator output rel err {err:.3e}"


@pytest.mark.parametrize("variant,T", [("default", 750), ("wavegan", 3000), ("unet", 750)])
def test_generator_long_sequence(variant, T):
    """Generator is length-agnostic (phase3/test.py:49: 30 s = 750 frames; BASELINE configs[4]: WaveGAN-style
    encoder on long sequences, here 2 minutes = 3000 frames in one call)."""
    cfg = O.make_cfg(**VARIANTS[variant])
    gen, _ = build(cfg)
    g = torch.Generator().manual_seed(22)
    audio = (torch.rand(1, T * 640, generator=g) - 0.5) * 0.6
    noise = torch.randn(1, T, cfg["noise_size"], generator=g)
    from music2dance_b200.utils import slice_audio_batch
    sl = O.slice_audio_batch(audio, 3200, 640, 2560)
    assert sl.shape[1] == T
    gen.eval()
    with torch.no_grad():
        out = gen(slice_audio_batch(audio.to(DEV), 3200, 640, 2560), [T], noise=noise.to(DEV))
    ref = O.generator_forward(oracle_params(gen), cfg, sl, noise, train=False)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < TOL_FP32


@pytest.mark.parametrize("variant,B,nc", [("default", 2, 2), ("default", 7, 3), ("wavegan", 3, 2), ("ablated", 4, 2),
                                          ("tanh", 2, 2)])
@pytest.mark.parametrize("graphs", [False, True])
def test_fused_trainer_vs_oracle(variant, B, nc, graphs):
    """Phase3Trainer (fused step, Adam included, optionally CUDA-graph replayed) against the
    oracle's train_step on the same staged batches.  Losses are continuous in the weights, so
    they are held to the north-star 1e-3 through `nc` chained Adam updates."""
    from music2dance_b200.trainer import Phase3Trainer
    cfg = O.make_cfg(n_critic_steps=nc, **VARIANTS[variant])
    gen, critic = build(cfg)
    G, D = oracle_params(gen), oracle_params(critic)
    tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=graphs)
    ad, ag = O.AdamState(D, cfg["lr_critic"]), O.AdamState(G, cfg["lr_gen"])
    for step in range(2):
        batches = [O.synthetic_batch(cfg, B, 4000 + step * nc + i) for i in range(nc)]
        tr.load_batches(torch.stack([b[0] for b in batches]), torch.stack([b[1] for b in batches]),
                        torch.stack([b[2] for b in batches]), torch.stack([b[3] for b in batches]),
                        batches[-1][4])
        tr.train_step()
        logs = tr.logs()
        for i, b in enumerate(batches):
            o = O.critic_iteration(G, D, cfg, b[0], b[1], b[2], b[3], ad)
            first = step == 0 and i == 0                        # identical state: north-star tolerance
            for k in ("loss_critic", "gp", "w_dist"):
                scalar_check(logs["critic"][i][k], o[k], TOL_NORTH_STAR if first else TOL_CHAINED,
                             f"step{step} it{i} {k}")
        b = batches[-1]
        o = O.generator_update(G, D, cfg, b[0], b[1], b[4], ag)
        for k in ("loss_gen", "l1", "tv"):
            scalar_check(logs["gen"][k], o[k], TOL_CHAINED, f"step{step} gen {k}")
        f = tr.fake_g.view(B, cfg["stick_length"], cfg["output_size"]).permute(0, 2, 1).cpu()
        assert float((f - o["fake"]).abs().max() / o["fake"].abs().max()) < TOL_CHAINED
    # parameters after 2*nc critic and 2 generator Adam steps: mean deviation in units of lr
    skip = set(O.pre_bn_bias_names(G))
    for mod, P, lr, steps in ((critic, D, cfg["lr_critic"], 2 * nc), (gen, G, cfg["lr_gen"], 2)):
        for k, v in mod.state_dict().items():
            if k in skip:
                continue
            if not v.is_floating_point():
                assert int(v) == int(P[k]), k
                continue
            d = (v.cpu() - P[k]).abs()
            if "running_" in k:
                assert float(d.max()) < TOL_DRIFT_BUF * max(float(P[k].abs().max()), 0.1), (k, float(d.max()))
            else:
                assert float(d.mean()) < TOL_DRIFT_LR * lr * steps + 1e-6 * float(P[k].abs().max()), (k, float(d.mean()))


def test_stock_gradient_penalty_fails_with_clear_message():
    """SURVEY §8(b) autograd contract: double backward through the drop-in critic is bypassed by the fused
    gradient_penalty; asking autograd for it (the reference's own losses.py:40-44 recipe) must say so."""
    cfg = O.make_cfg(ablated=True)
    _, critic = build(cfg)
    x = torch.rand(2, cfg["output_size"], cfg["stick_length"], device=DEV, requires_grad=True)
    out = critic(x)
    with pytest.raises(RuntimeError, match="music2dance_b200.losses.gradient_penalty"):
        torch.autograd.grad(outputs=out, inputs=x, grad_outputs=torch.ones_like(out), create_graph=True,
                            retain_graph=True, only_inputs=True)
    g, = torch.autograd.grad(outputs=critic(x), inputs=x, grad_outputs=torch.ones_like(out))   # first order still works
    assert g.shape == x.shape and bool(torch.isfinite(g).all())


def _sync_trainer_from_oracle(tr, gen, critic, G, D, ad, ag):
    """Copy the oracle's state (parameters, BatchNorm buffers, Adam moments and step counts) into the trainer: after
    this both sides start the next iteration from IDENTICAL state, so every iteration is a single-iteration
    comparison (the protocol of oracle/validate_vs_reference.py:93-101)."""
    torch.cuda.synchronize()
    gen.load_state_dict({k: v.clone() for k, v in G.items()}, strict=True)
    critic.load_state_dict({k: v.clone() for k, v in D.items()}, strict=True)
    for eng, st, m, v, tabs in ((tr.de, ad, tr.mD, tr.vD, [tr.apD, tr.apD_late]), (tr.ge, ag, tr.mG, tr.vG, [tr.apG])):
        base = eng.fp.flat.data_ptr()
        for name, prm in eng.fp.params.items():
            if name not in st.m:
                continue
            off = (prm.data_ptr() - base) // 4
            if off >= m.numel():
                continue                                      # dead LinearBlock branch (Q1): never updated
            m[off:off + prm.numel()].copy_(st.m[name].reshape(-1))
            v[off:off + prm.numel()].copy_(st.v[name].reshape(-1))
        t = max(st.t.values()) if st.t else 0
        for tab in tabs:
            if tab is not None:
                tab.counters[0] = t
    tr._ensure_packed()
    torch.cuda.synchronize()


@pytest.mark.parametrize("variant,B,nc", [("default", 7, 8), ("tv", 2, 2), ("noise_enhanced", 2, 2), ("unet", 2, 2),
                                          ("wavegan", 2, 2), ("ablated", 3, 3), ("tanh", 2, 2), ("relu", 2, 2)])
def test_trainer_every_iteration_vs_oracle_resynced(variant, B, nc):
    """North-star tolerance (1e-3 relative on losses and gradient-penalty terms) on EVERY critic iteration and on the
    generator update of a train step — including the BASELINE configuration itself (default.yaml: batch 7,
    n_critic 8), tv.yaml, noise_enhanced.yaml, the U-Net / WaveGAN encoders and the three code activations.  After
    each comparison the oracle's post-update state is copied into the trainer, so chained Adam steps cannot amplify a
    ReLU / L1 kink flip into the next comparison (tests/parity.py: TOL_CHAINED is what the un-synced test needs)."""
    from music2dance_b200.trainer import Phase3Trainer
    cfg = O.make_cfg(n_critic_steps=nc, **VARIANTS[variant])
    gen, critic = build(cfg)
    G, D = oracle_params(gen), oracle_params(critic)
    tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=False)
    ad, ag = O.AdamState(D, cfg["lr_critic"]), O.AdamState(G, cfg["lr_gen"])
    T, Oo = cfg["stick_length"], cfg["output_size"]
    for step in range(2 if nc < 8 else 1):
        batches = [O.synthetic_batch(cfg, B, 5000 + step * nc + i) for i in range(nc)]
        tr.load_batches(torch.stack([b[0] for b in batches]), torch.stack([b[1] for b in batches]),
                        torch.stack([b[2] for b in batches]), torch.stack([b[3] for b in batches]),
                        batches[-1][4])
        for i, b in enumerate(batches):
            with torch.cuda.device(tr.dev):
                tr.critic_iteration(i)                            # generator forward, critic backward, Adam, re-layout
            torch.cuda.synchronize()
            got = dict(zip(("loss_critic", "gp", "w_dist"), tr.log_c[i, :3].tolist()))
            o = O.critic_iteration(G, D, cfg, b[0], b[1], b[2], b[3], ad)
            for k in ("loss_critic", "gp", "w_dist"):
                scalar_check(got[k], o[k], TOL_NORTH_STAR, f"{variant} step{step} it{i} {k}")
            f = tr.fake_c[i].view(B, T, Oo).permute(0, 2, 1).cpu()
            assert float((f - o["fake"]).abs().max() / o["fake"].abs().max()) < TOL_NORTH_STAR, f"it{i} generated poses"
            _sync_trainer_from_oracle(tr, gen, critic, G, D, ad, ag)
        b = batches[-1]
        with torch.cuda.device(tr.dev):
            tr.generator_update()
        torch.cuda.synchronize()
        got = dict(zip(("loss_gen", "l1", "tv"), tr.log_g[:3].tolist()))
        o = O.generator_update(G, D, cfg, b[0], b[1], b[4], ag)
        for k in ("loss_gen", "l1", "tv"):
            scalar_check(got[k], o[k], TOL_NORTH_STAR, f"{variant} step{step} gen {k}")
        _sync_trainer_from_oracle(tr, gen, critic, G, D, ad, ag)


@pytest.mark.parametrize("variant,groups", [("default", "4"), ("default", "1,3"), ("unet", "2,2"), ("wavegan", "4"),
                                            ("noise_enhanced", "1,1,2")])
def test_batched_generator_passes_match_one_pass_per_iteration(variant, groups, monkeypatch):
    """trainer.gen_groups: the generator forwards of several critic iterations as ONE pass (per-batch BatchNorm
    statistics, running statistics advanced batch by batch) against one pass per iteration — generated poses of every
    iteration, the generator's BatchNorm buffers after the step, and the step's logs."""
    from music2dance_b200.trainer import Phase3Trainer
    nc, B = 4, 2
    cfg = O.make_cfg(n_critic_steps=nc, **VARIANTS[variant])
    res = []
    for spec in (",".join(["1"] * nc), groups):
        monkeypatch.setenv("M2D_GEN_GROUPS", spec)
        gen, critic = build(cfg)
        tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=False)
        assert [tr.gen_groups[i] for i in sorted(tr.gen_groups)] == [int(x) for x in spec.split(",")]
        bs = [O.synthetic_batch(cfg, B, 9100 + i) for i in range(nc)]
        tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])
        tr.train_step()
        logs = tr.logs()
        bufs = {k: v.detach().cpu().clone() for k, v in gen.named_buffers()}
        res.append((tr.fake_c.cpu().clone(), bufs, logs))
    (f0, b0, l0), (f1, b1, l1) = res
    for i in range(nc):
        e = float((f1[i] - f0[i]).abs().max() / f0[i].abs().max())
        assert e < 2e-5, f"{variant} {groups}: generated poses of iteration {i} differ by {e:.2e}"
    for k, v in b0.items():
        if v.is_floating_point():
            e = float((b1[k] - v).abs().max() / max(float(v.abs().max()), 0.1))
            assert e < 1e-5, f"{variant} {groups}: {k} differs by {e:.2e}"
        else:
            assert torch.equal(b1[k], v), k                       # num_batches_tracked: nc + 1 forwards either way
    for i in range(nc):
        for k in ("loss_critic", "gp", "w_dist"):
            scalar_check(l1["critic"][i][k], l0["critic"][i][k], TOL_NORTH_STAR if i == 0 else TOL_CHAINED,
                         f"{variant} {groups} it{i} {k}")
    for k in ("loss_gen", "l1"):
        scalar_check(l1["gen"][k], l0["gen"][k], TOL_CHAINED, f"{variant} {groups} gen {k}")


@pytest.mark.parametrize("variant,B", [("default", 7), ("ablated", 4), ("tanh", 2)])
def test_critic_gradients_full_tensor_vs_oracle(variant, B):
    """Every critic gradient of one iteration compared over ALL its elements (the fixture checks are 64-sample digests):
    relative l2 error against the oracle's autograd gradients on the same batch.  A ReLU unit whose pre-activation sits
    within rounding noise of zero moves an l2 norm by ~1/sqrt(units), hence the kink bound rather than 1e-3 on l2;
    bias gradients are cancellation-dominated sums (tests/parity.py: TOL_GRAD_BIAS).  Measured on B200 (worst tensor):
    default B = 7: weights 1.2e-3, biases 2.5e-3; ablated B = 4: 6e-6 / 3e-6; tanh B = 2: 3.3e-3 / 3.4e-3."""
    from music2dance_b200.trainer import Phase3Trainer
    cfg = O.make_cfg(n_critic_steps=1, **VARIANTS[variant])
    gen, critic = build(cfg)
    G, D = oracle_params(gen), oracle_params(critic)
    tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=False)
    b = O.synthetic_batch(cfg, B, 7100)
    tr.load_batches(b[0][None], b[1][None], b[2][None], b[3][None], b[4])
    with torch.cuda.device(tr.dev):
        tr.critic_iteration(0, update=False)            # gradients unpacked into the parameter layout
    torch.cuda.synchronize()
    o = O.critic_iteration(G, D, cfg, b[0], b[1], b[2], b[3], None)
    gmax = max(float(g.abs().max()) for g in o["grads"].values() if g is not None)
    worst_w = worst_b = 0.0
    for name, g_ref in o["grads"].items():
        got = tr.de.fp.G[name].detach().cpu()
        if g_ref is None or float(g_ref.norm()) < 1e-7 * gmax:        # fc2.bias: the two Wasserstein terms cancel exactly
            assert float(got.abs().max()) < 1e-5 * gmax, (name, float(got.abs().max()))
            continue
        e = float((got - g_ref).norm() / g_ref.norm())
        if name.endswith(".bias"):
            worst_b = max(worst_b, e)
            assert e < TOL_GRAD_BIAS, f"{variant} {name}: l2 error {e:.3e}"
        else:
            worst_w = max(worst_w, e)
            assert e < TOL_KINK_L2, f"{variant} {name}: l2 error {e:.3e}"
    print(f"[{variant} B={B}] worst relative l2 error: weights {worst_w:.2e}, biases {worst_b:.2e}")


def test_per_iteration_graphs_match_single_graph():
    """The multi-GPU step structure (one CUDA graph per critic iteration, optimiser / re-layout graphs in between,
    generator forwards pipelined one iteration ahead, audio_d.l5 / l6 re-layout forked into the next graph) run on ONE
    GPU must reproduce the single-graph step: same kernels, same order per stream -> same losses to fp32 noise."""
    from music2dance_b200.trainer import Phase3Trainer
    cfg = O.make_cfg(n_critic_steps=3)
    B = 2
    logs = []
    for per_iter in (False, True):
        gen, critic = build(cfg)
        tr = Phase3Trainer(gen, critic, cfg, B, use_graphs=True, per_iteration_graphs=per_iter)
        out = []
        for step in range(3):
            bs = [O.synthetic_batch(cfg, B, 7000 + step * 3 + i) for i in range(3)]
            tr.load_batches(*[torch.stack([b[j] for b in bs]) for j in range(4)], bs[-1][4])
            tr.train_step()
            out.append(tr.logs())
        logs.append(out)
    for a, b in zip(*logs):
        for ca, cb in zip(a["critic"], b["critic"]):
            for k in ("loss_critic", "gp", "w_dist"):
                scalar_check(cb[k], ca[k], 1e-3, f"per-iteration graphs {k}")
        scalar_check(b["gen"]["loss_gen"], a["gen"]["loss_gen"], 1e-3, "per-iteration graphs loss_gen")


def test_critic_batch_additivity():
    """Data-parallel property (no BatchNorm in the critic): the critic gradients of a batch
    equal the mean of the gradients of its two halves — what the all-reduce relies on."""
    from music2dance_b200.trainer import Phase3Trainer
    cfg = O.make_cfg(n_critic_steps=1)
    B = 4
    batch = O.synthetic_batch(cfg, B, 4100)
    grads = []
    for sl in (slice(0, 4), slice(0, 2), slice(2, 4)):
        gen, critic = build(cfg)
        n = sl.stop - sl.start
        tr = Phase3Trainer(gen, critic, cfg, n, use_graphs=False)
        # the generator's BatchNorm couples the batch, so feed every run the SAME fake poses
        tr.load_batches(batch[0][sl][None], batch[1][sl][None], batch[2][sl][None], batch[3][sl][None], batch[4][sl])
        if n == B:
            full = tr
            tr.critic_iteration(0, update=False)
            fake_full = tr.fake_c.clone().view(B, -1)
        else:
            orig = tr.G.forward

            def fwd(audio, noise, Bn, T, train=True, out=None, _sl=sl, _o=orig):
                r = _o(audio, noise, Bn, T, train=train, out=out)
                out.t.view(Bn, -1).copy_(fake_full[_sl])
                return r
            tr.G.forward = fwd
            tr.critic_iteration(0, update=False)
        grads.append(tr.de.fp.grad[:tr.de.fp.n_live_padded].clone())
    mean = 0.5 * (grads[1] + grads[2])
    e = float((grads[0] - mean).abs().max() / grads[0].abs().max())
    assert e < TOL_GRAD, f"batch additivity violated: {e:.3e}"
