"""Input pipeline: the CPU oracle against golden batches produced by the reference's own
SequenceDataset / collate_fn / WeightedRandomSampler / DataLoader, and (GPU) the device-resident loader
against the oracle and the same fixture.  Indexing work: everything is compared BIT-EXACT."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as DO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phase3_data.npz")


def _epochs(gold):
    data, cfg = DO.synthetic_dataset(), DO.dataset_config()
    torch.manual_seed(int(gold["seed_torch"]))
    np.random.seed(int(gold["seed_numpy"]))
    return data, cfg, DO.class_weights(data["labels"]), int(gold["batch"])


def _check_batch(gold, n, real, lengths, audio, labels, dirs):
    assert np.array_equal(real.cpu().numpy(), gold[f"b{n}/real"]), f"batch {n}: poses differ"
    a = audio.cpu()
    assert np.array_equal(a[:, :64].numpy(), gold[f"b{n}/audio_head"]), f"batch {n}: audio head differs"
    assert np.array_equal(a[:, -64:].numpy(), gold[f"b{n}/audio_tail"]), f"batch {n}: audio tail differs"
    dg = gold[f"b{n}/audio_digest"]
    assert float(a.double().sum()) == dg[0] and float(a.double().abs().max()) == dg[1], f"batch {n}: audio digest"
    assert np.array_equal(labels.cpu().numpy(), gold[f"b{n}/labels"])
    assert list(lengths) == gold[f"b{n}/lengths"].tolist()
    assert list(dirs) == gold[f"b{n}/dirs"].tolist()


def test_oracle_matches_reference_loader_fixture():
    gold = np.load(GOLD)
    data, cfg, w, bs = _epochs(gold)
    n = 0
    for ep in range(2):
        for (real, lengths, audio, labels, dirs), _, _ in DO.epoch(data, cfg, bs, w):
            _check_batch(gold, n, real, lengths, audio, labels, dirs)
            n += 1
    assert n == int(gold["n_batches"])


def test_ragged_last_batch_and_bounds():
    """9 sequences at batch 4: the last batch of an epoch has one entry (the reference's loader yields it too)."""
    gold = np.load(GOLD)
    assert [gold[f"b{n}/real"].shape[0] for n in range(int(gold["n_batches"]))] == [4, 4, 1, 4, 4, 1]


@pytest.mark.gpu
def test_device_loader_bit_exact_vs_reference_fixture_and_oracle():
    from music2dance_b200.data import DeviceSequenceDataset, Phase3Loader
    gold = np.load(GOLD)
    data, cfg, w, bs = _epochs(gold)
    ds = DeviceSequenceDataset(data, cfg, "cuda:0")
    assert (ds.stick_length, ds.audio_length, ds.ratio) == (120, 76800, 640)
    loader = Phase3Loader(ds, bs, w)
    n = 0
    for ep in range(2):
        for real, lengths, audio, labels, dirs in loader:
            assert real.is_cuda and audio.is_cuda and real.shape[1:] == (120, 23, 3)
            _check_batch(gold, n, real, lengths, audio, labels, dirs)
            n += 1
    assert n == int(gold["n_batches"])
    # same seeds again: the oracle and the device loader consume the generators identically
    torch.manual_seed(11)
    np.random.seed(12)
    ref = [(b[0].clone(), b[2].clone(), bi, st) for b, bi, st in DO.epoch(data, cfg, 3, w)]
    torch.manual_seed(11)
    np.random.seed(12)
    for (r, a, _, _), (real, _, audio, _, _) in zip(ref, Phase3Loader(ds, 3, w)):
        assert torch.equal(real.cpu(), r) and torch.equal(audio.cpu(), a)


@pytest.mark.gpu
def test_crop_edges_and_errors():
    """First / last admissible start of every sequence; an out-of-range crop raises instead of reading past the data."""
    from music2dance_b200.data import DeviceSequenceDataset
    data, cfg = DO.synthetic_dataset(), DO.dataset_config()
    ds = DeviceSequenceDataset(data, cfg, "cuda:0")
    idx = list(range(len(ds)))
    for starts in ([0] * len(ds), [ds.lengths[i] - 120 for i in idx]):
        real, audio = ds.crop(idx, starts)
        for b, (i, s) in enumerate(zip(idx, starts)):
            assert np.array_equal(real[b].cpu().numpy(), data["sequences"][i][s:s + 120].astype(np.float32))
            assert np.array_equal(audio[b].cpu().numpy(), data["musics"][i][s * 640:s * 640 + 76800])
    with pytest.raises(IndexError):
        ds.crop([0], [ds.lengths[0] - 119])
