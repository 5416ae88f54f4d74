"""Run-directory, checkpoint and logging formats either side of the train step, so that a run of
``Phase3Trainer`` is interchangeable with the reference's tooling (``phase3/test.py`` loads ``gpgen_*.pt`` with
``load_state_dict``; tensorboard reads the scalar tags below).

Mirrors ``phase3/train.py``: run directory naming and sub-folders (:37-42,107-111), the train/val/test split and
``trainvaltest_samples.json`` (:113-131), ``model_gen.txt`` / ``model_critic.txt`` (:173-178), the tensorboard
scalar tags and their sign conventions (:239-243,261) and the checkpoint cadence / file names (:267-276).
Pure host code: nothing here touches the GPU except ``torch.save`` reading the parameters.
"""
from __future__ import annotations

import datetime
import json
import os
from collections import OrderedDict

import numpy as np
import torch

VALIDATION_SPLIT, TEST_SPLIT, SPLIT_SEED = 0.2, 0.5, 14           # train.py:113-115


def run_name(name, now=None):
    """train.py:40: ``%Y%m%d-%H%M%S_<name>``."""
    now = datetime.datetime.now() if now is None else now
    return now.strftime("%Y%m%d-%H%M%S") + "_" + name


def make_run_dirs(name, root="./runs", now=None):
    """train.py:38-42,107-111 -> dict(run, logging, samples, models).  Like the reference, an existing run
    directory is an error (``os.makedirs`` without exist_ok)."""
    if not os.path.exists(root):
        os.makedirs(root)
    run = os.path.join(root, run_name(name, now))
    os.makedirs(run)
    dirs = dict(run=run, logging=os.path.join(run, "logging"), samples=os.path.join(run, "samples"),
                models=os.path.join(run, "models"))
    os.makedirs(dirs["samples"])
    os.makedirs(dirs["models"])
    return dirs


def split_indices(dataset_size, validation_split=VALIDATION_SPLIT, test_split=TEST_SPLIT, random_seed=SPLIT_SEED):
    """train.py:116-125: shuffle with numpy's legacy generator seeded 14 (``np.random.seed`` + ``np.random.shuffle``
    == ``RandomState(seed).shuffle``; a private generator keeps the caller's global numpy state untouched), then
    test = first tsplit, validation = [tsplit, vsplit), train = the rest.  Bit-exact index lists."""
    indices = list(range(dataset_size))
    vsplit = int(np.floor(validation_split * dataset_size))
    tsplit = int(np.floor(test_split * vsplit))
    np.random.RandomState(random_seed).shuffle(indices)
    return indices[vsplit:], indices[tsplit:vsplit], indices[:tsplit]


def write_samples_json(run_dir, dirs, train_indices, val_indices, test_indices):
    """train.py:126-131: ``trainvaltest_samples.json`` with the sequence directory names of each split."""
    samples = {"train_samples": [dirs[i] for i in train_indices],
               "val_samples": [dirs[i] for i in val_indices],
               "test_samples": [dirs[i] for i in test_indices]}
    path = os.path.join(run_dir, "trainvaltest_samples.json")
    with open(path, "w+") as f:
        json.dump(samples, f)
    return path


def write_model_descriptions(run_dir, gen, critic):
    """train.py:173-178: ``str(module)`` of both networks (the drop-in modules keep the reference's module tree, so
    the text is the reference's up to the class repr of the parameter containers)."""
    for fname, m in (("model_gen.txt", gen), ("model_critic.txt", critic)):
        with open(os.path.join(run_dir, fname), "w+") as f:
            f.write(str(m))


def train_scalars(logs, err_l1=None):
    """train.py:239-243 — tag -> value with the reference's sign flips (``loss_critic`` and ``w_dist`` are logged
    negated).  `logs` = what ``Phase3Trainer.logs()`` returns: ``{'critic': [one dict per critic iteration],
    'gen': {...}}``; the reference logs on the iteration where ``total_iterations % n_critic_steps == 0``, i.e. the
    LAST critic iteration of the train step (train.py:218-219), together with that step's generator update.  A flat
    dict with the five keys is accepted too."""
    if "critic" in logs and "gen" in logs:
        c, g = logs["critic"][-1], logs["gen"]
        logs = {"loss_critic": c["loss_critic"], "gp": c["gp"], "w_dist": c["w_dist"],
                "loss_gen": g["loss_gen"], "l1": g["l1"]}
    out = OrderedDict()
    out["loss_critic"] = -float(logs["loss_critic"])
    out["loss_gen"] = float(logs["loss_gen"])
    out["gp"] = float(logs["gp"])
    out["w_dist"] = -float(logs["w_dist"])
    out["l1_loss_train"] = float(logs["l1"] if err_l1 is None else err_l1)
    return out


def log_train_scalars(writer, logs, total_iterations, err_l1=None):
    """``writer`` = anything with ``add_scalar(tag, value, step)`` (``torch.utils.tensorboard.SummaryWriter`` in the
    reference, train.py:106)."""
    for tag, v in train_scalars(logs, err_l1).items():
        writer.add_scalar(tag, v, total_iterations)


def log_val_scalar(writer, e_val_loss, total_iterations):
    """train.py:260-261: mean of the per-batch validation L1 losses."""
    writer.add_scalar("l1_loss_val", float(e_val_loss), total_iterations)


def checkpoints_due(epoch):
    """train.py:267-276 for the 0-based ``epoch`` just finished -> list of ("gen" | "critic", file name).
    Generator every 100 epochs up to 1000, generator + critic every 5000 (so epoch 999 -> one generator file; the
    two rules never overlap)."""
    e, out = epoch + 1, []
    if e <= 1000 and e % 100 == 0:
        out.append(("gen", "gpgen_{}.pt".format(e)))
    if e % 5000 == 0:
        out.append(("gen", "gpgen_{}.pt".format(e)))
        out.append(("critic", "gpcritic_{}.pt".format(e)))
    return out


def save_checkpoints(gen, critic, model_dir, epoch):
    """``torch.save(module.state_dict(), models/gp{gen,critic}_<epoch+1>.pt)`` (train.py:268-276).  The drop-in
    modules keep the reference's state_dict keys / shapes / order, so these files load into the
    reference's modules with ``strict=True`` and vice versa.  Tensors are saved from their current device like the
    reference does.  Returns the paths written."""
    written = []
    for which, fname in checkpoints_due(epoch):
        path = os.path.join(model_dir, fname)
        torch.save((gen if which == "gen" else critic).state_dict(), path)
        written.append(path)
    return written


def load_generator(gen, path, map_location=None):
    """phase3/test.py:67: ``model.load_state_dict(torch.load(model_weights, map_location='cpu'))`` (strict)."""
    sd = torch.load(path, map_location=map_location)
    gen.load_state_dict(sd, strict=True)
    return gen
