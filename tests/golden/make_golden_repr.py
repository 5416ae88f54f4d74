"""`str(module)` of the UNMODIFIED reference generator / critic for every config variant (what phase3/train.py:173-178
writes to model_gen.txt / model_critic.txt) -> tests/golden/module_repr.json.  Build container only (/root/reference).
    python tests/golden/make_golden_repr.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import phase3_oracle as O          # noqa: E402
from oracle import reference_harness as R      # noqa: E402
from tests.parity import VARIANTS              # noqa: E402

out = {}
for name, over in VARIANTS.items():
    cfg = O.make_cfg(**over)
    gen, critic = R.build_models(cfg)
    out[name] = {"gen": str(gen), "critic": str(critic)}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "module_repr.json"), "w") as f:
    json.dump(out, f, indent=0)
print({k: (len(v["gen"]), len(v["critic"])) for k, v in out.items()})
