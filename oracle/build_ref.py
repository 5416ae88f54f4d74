"""Recipe: stage the UNMODIFIED reference modules of the phase3 hot path under oracle/_ref/ (git-ignored, travels to
the GPU box with the snapshot) so that `bench.py --impl reference` and the `torch_eager_gpu` leg drive the reference's
own stock code where /root/reference does not exist.

    python oracle/build_ref.py            # copies phase3/archis/default.py, losses.py, utils.py + MANIFEST.json

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under music2dance_b200/ imports these files.  The files are byte-for-byte
copies (sha256 recorded in oracle/_ref/MANIFEST.json); they are never committed (.gitignore: oracle/_ref/).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["phase3/archis/default.py", "losses.py", "utils.py"]


def build(src=None, verbose=False):
    src = src or os.environ.get("M2D_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(src, "phase3", "archis")):
        return None                                   # GPU box: the prebuilt copy (if any) is used as it is
    man = {"source": src, "files": {}}
    for rel in FILES:
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
        with open(d, "rb") as f:
            man["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    if verbose:
        print(json.dumps(man, indent=1))
    return DST


def available():
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    p = build(verbose=True)
    sys.exit(0 if p else 1)
