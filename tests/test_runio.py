"""Run directory / checkpoint / logging formats (music2dance_b200/runio.py) against the statements of
phase3/train.py they mirror (:37-42,107-131,173-178,239-243,261,267-276).  Host-only: runs without a GPU; the
state_dict round trip goes through the REFERENCE-shaped module tree on CPU (parameter containers only — no forward)."""
import datetime
import json
import os

import numpy as np
import pytest
import torch

from music2dance_b200 import runio


def test_split_is_the_reference_shuffle():
    # train.py:116-125 restated with the GLOBAL numpy generator exactly as the reference writes it
    for n in (0, 1, 7, 53, 1000):
        indices = list(range(n))
        vsplit = int(np.floor(.2 * n))
        tsplit = int(np.floor(.5 * vsplit))
        np.random.seed(14)
        np.random.shuffle(indices)
        tr, va, te = runio.split_indices(n)
        assert tr == indices[vsplit:] and va == indices[tsplit:vsplit] and te == indices[:tsplit]
        assert sorted(tr + va + te) == list(range(n))


def test_split_leaves_global_numpy_state_alone():
    np.random.seed(3)
    a = np.random.rand()
    np.random.seed(3)
    runio.split_indices(100)
    assert np.random.rand() == a


def test_run_dirs_and_samples_json(tmp_path):
    now = datetime.datetime(2019, 7, 4, 13, 5, 9)
    assert runio.run_name("exp", now) == "20190704-130509_exp"
    d = runio.make_run_dirs("exp", root=str(tmp_path / "runs"), now=now)
    assert os.path.isdir(d["samples"]) and os.path.isdir(d["models"])
    assert d["run"].endswith("runs/20190704-130509_exp") and d["logging"].endswith("/logging")
    names = [f"seq_{i:03d}" for i in range(20)]
    tr, va, te = runio.split_indices(len(names))
    path = runio.write_samples_json(d["run"], names, tr, va, te)
    assert os.path.basename(path) == "trainvaltest_samples.json"
    js = json.load(open(path))
    assert list(js) == ["train_samples", "val_samples", "test_samples"]
    assert (len(js["train_samples"]), len(js["val_samples"]), len(js["test_samples"])) == (16, 2, 2)
    assert js["test_samples"] == [names[i] for i in te]


def test_checkpoint_cadence_matches_train_py():
    # train.py:267-276 restated literally
    for epoch in list(range(0, 1300)) + [4998, 4999, 5000, 9999, 14999]:
        want = []
        if (epoch + 1) <= 1000 and (epoch + 1) % 100 == 0:
            want.append(("gen", "gpgen_{}.pt".format(epoch + 1)))
        if (epoch + 1) % 5000 == 0:
            want.append(("gen", "gpgen_{}.pt".format(epoch + 1)))
            want.append(("critic", "gpcritic_{}.pt".format(epoch + 1)))
        assert runio.checkpoints_due(epoch) == want
    assert runio.checkpoints_due(99) == [("gen", "gpgen_100.pt")]
    assert runio.checkpoints_due(100) == []
    assert runio.checkpoints_due(4999) == [("gen", "gpgen_5000.pt"), ("critic", "gpcritic_5000.pt")]


class _Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, v, step):
        self.rows.append((tag, v, step))


def test_scalar_tags_and_signs():
    logs = dict(loss_critic=-3.5, gp=0.25, w_dist=-1.5, loss_gen=12.0, l1=0.75)
    w = _Writer()
    runio.log_train_scalars(w, logs, 40)
    assert w.rows == [("loss_critic", 3.5, 40), ("loss_gen", 12.0, 40), ("gp", 0.25, 40), ("w_dist", 1.5, 40),
                      ("l1_loss_train", 0.75, 40)]
    runio.log_val_scalar(w, np.mean([0.5, 1.5]), 40)
    assert w.rows[-1] == ("l1_loss_val", 1.0, 40)


def test_scalars_from_trainer_logs_structure():
    """`Phase3Trainer.logs()` returns {'critic': [nc dicts], 'gen': {...}} (trainer.LOG_CRITIC / LOG_GEN): the scalars
    the reference writes at train.py:239-243 come from the LAST critic iteration and the generator update."""
    from music2dance_b200.trainer import LOG_CRITIC, LOG_GEN
    nc = 8
    crit = [dict(zip(LOG_CRITIC, [-(i + 1.0), 0.1 * i, -0.5 * i, 1.0, 2.0])) for i in range(nc)]
    gen = dict(zip(LOG_GEN, [12.0, 0.75, 0.0, 1.0, 2.0]))
    w = _Writer()
    runio.log_train_scalars(w, {"critic": crit, "gen": gen}, 8)
    assert w.rows == [("loss_critic", 8.0, 8), ("loss_gen", 12.0, 8), ("gp", 0.1 * 7, 8), ("w_dist", 3.5, 8),
                      ("l1_loss_train", 0.75, 8)]


@pytest.mark.parametrize("variant", ["default", "wavegan", "unet", "ablated", "tanh", "relu", "tv", "noise_enhanced"])
def test_module_text_matches_reference(variant):
    """phase3/train.py:173-178 writes `str(gen)` / `str(critic)` to model_gen.txt / model_critic.txt: the drop-in
    modules print EXACTLY what the reference's modules print (fixture: tests/golden/make_golden_repr.py ran the
    unmodified reference) — structural nodes named after the reference's classes, parameter-free layers included."""
    import json
    from music2dance_b200.archis.default import (AblatedSequenceDiscriminator, SequenceDiscriminator,
                                                 SequenceGenerator)
    from oracle import phase3_oracle as O
    from tests.parity import VARIANTS
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "module_repr.json")))[variant]
    cfg = O.make_cfg(**VARIANTS[variant])
    gen = SequenceGenerator(cfg["audio_feat_samples"], cfg["input_vector_size"], cfg["latent_vector_size"], cfg["size"],
                            cfg["output_size"], cfg["noise_size"], cfg["nblocks_gen"], cfg["n_cells"], cfg["enc_type"],
                            cfg["activ"], "cpu")
    cls = AblatedSequenceDiscriminator if cfg["ablated"] else SequenceDiscriminator
    critic = cls(cfg["output_size"], cfg["channels"], cfg["code_size"], cfg["stick_length"],
                 init_ker=cfg["init_kernel"], activ=cfg["activ"], device="cpu")
    assert str(gen) == gold["gen"]
    assert str(critic) == gold["critic"]


def test_checkpoint_files_round_trip_with_reference_keys(tmp_path):
    """Files written at the reference cadence hold the drop-in modules' state_dict (= the reference's keys, SURVEY
    Appendix A) and load back strictly."""
    from music2dance_b200.archis.default import SequenceDiscriminator, SequenceGenerator
    torch.manual_seed(0)
    gen = SequenceGenerator(3200, 250, 240, 256, 69, 10, 2, 3, "default", "id", "cpu")
    critic = SequenceDiscriminator(69, 128, 100, 120, init_ker=25, activ="id", device="cpu")
    assert runio.save_checkpoints(gen, critic, str(tmp_path), 5) == []
    (p,) = runio.save_checkpoints(gen, critic, str(tmp_path), 99)
    assert os.path.basename(p) == "gpgen_100.pt"
    sd = torch.load(p)
    assert list(sd) == list(gen.state_dict())
    # the keys are the REFERENCE module's (fixture written from the unmodified reference generator)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase3_default.npz"))
    ref_keys = {k[len("init/init/"):-len("/sum")] for k in gold.files if k.startswith("init/init/") and k.endswith("/sum")}
    assert set(sd) <= ref_keys and len(sd) > 50
    torch.manual_seed(1)
    gen2 = SequenceGenerator(3200, 250, 240, 256, 69, 10, 2, 3, "default", "id", "cpu")
    runio.load_generator(gen2, p, map_location="cpu")
    for (k, a), (_, b) in zip(gen.state_dict().items(), gen2.state_dict().items()):
        assert torch.equal(a, b), k
    paths = runio.save_checkpoints(gen, critic, str(tmp_path), 4999)
    assert [os.path.basename(x) for x in paths] == ["gpgen_5000.pt", "gpcritic_5000.pt"]
    assert list(torch.load(paths[1])) == list(critic.state_dict())
    runio.write_model_descriptions(str(tmp_path), gen, critic)
    assert open(tmp_path / "model_gen.txt").read() == str(gen)
    assert open(tmp_path / "model_critic.txt").read().startswith("SequenceDiscriminator(\n  (stick_d): StickDiscriminator(")
