"""the algorithmic work per sequence (what `roofline.achieved` is built from) must be recomputable
from layer shapes.  tools/algorithmic_work.py derives it from the drop-in modules' own layer containers; the figures
below are the survey's hook-verified MAC counts of the REFERENCE modules, so this also checks that
the drop-in module tree has the reference's layer shapes.  Host-only."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("algorithmic_work", os.path.join(ROOT, "tools", "algorithmic_work.py"))
aw = importlib.util.module_from_spec(spec)
spec.loader.exec_module(aw)

SURVEY = {  # enc: (G MMAC, step GMAC, step GFLOP)
    "default": (1264.66, 69.92, 139.8),
    "wavegan": (3015.31, 89.17, 178.3),
    "unet": (8946.38, 154.4, 308.8),
}


@pytest.mark.parametrize("enc", ["default", "wavegan", "unet"])
def test_macs_from_layer_shapes_match_survey(enc):
    w = aw.work(enc)
    G, step, gflop = SURVEY[enc]
    assert abs(w["G_macs"] / 1e6 - G) < 0.006
    assert abs(w["A_macs"] / 1e6 - 1002.24) < 0.006            # critic audio branch (default.py:294-319)
    assert abs(w["S_macs"] / 1e6 - 83.08) < 0.006              # pose branch (default.py:322-346)
    assert abs(w["F_macs"] / 1e6 - 0.026) < 0.0005             # fusion MLP
    assert w["step_macs_per_sequence"] == 11 * w["G_macs"] + 49 * w["A_macs"] + 83 * w["S_macs"]
    assert abs(w["step_macs_per_sequence"] / 1e9 - step) < 0.06
    assert abs(w["step_gflop_per_sequence"] - gflop) < 0.06


def test_parameter_counts_and_hbm_figures():
    w = aw.work("default")
    assert w["critic_params"] == 10_436_041
    assert w["generator_live_params"] == 4_583_859              # the 132,608 dead fc1/bn1 parameters get no gradient
    assert w["generator_params"] - w["generator_live_params"] == 132_608
    assert w["adam_bytes_per_step"] == 28 * (8 * 10_436_041 + 4_583_859)          # 2.47 GB
    assert round(w["allreduce_bytes_per_step"] / 1e6) == 352
    assert abs(w["reference_executed_macs_per_sequence"] / 1e9 - 124.6) < 0.06
