"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/m2d.h declares;
the ctypes signature table covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "m2d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m2d_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    path = ge.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/m2d.h but not exported by libm2d_b200.so"


def test_ctypes_table_matches_header():
    from music2dance_b200 import _lib
    table = set(_lib.SIGNATURES) | {"m2d_last_error"}
    assert table == set(declared_symbols())


def test_struct_layouts_match_header():
    """Field order of the two argument structs must follow the header (ctypes mirrors it by hand)."""
    from music2dance_b200 import _lib
    src = open(os.path.join(ROOT, "include", "m2d.h")).read()
    for cname, cls in (("m2d_rowconv_args", _lib.RowConvArgs), ("m2d_wgrad_args", _lib.WgradArgs)):
        end = src.index("} " + cname + ";")
        body = src[src.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const float* x" / "long long x_bs" / "int N, T, Cc"
            names = re.sub(r"^(const\s+)?(float\s*\*|double\s*\*|long long|int|float)\s*", "", decl)
            fields += [n.strip().lstrip("*").strip() for n in names.split(",")]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "music2dance_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_cpu_inputs_fail_loudly():
    import pytest
    import torch
    from music2dance_b200 import losses, utils
    with pytest.raises(RuntimeError):
        losses.tv_loss(torch.zeros(1, 3, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            utils.slice_audio_batch(torch.zeros(2, 6400), 3200, 640, 2560)


def test_product_config_matches_oracle_config():
    """music2dance_b200/config.py (product side: bench.py, tools/) restates the same default.yaml constants and
    synthetic inputs as the oracle's; the product never imports oracle/, so the two are pinned to each other here."""
    import torch
    from music2dance_b200 import config as C
    from oracle import phase3_oracle as O
    for over in ({}, {"enc_type": "wavegan"}, {"ablated": True, "activ": "tanh"}):
        assert C.make_cfg(**over) == O.make_cfg(**over)
    a, b = C.synthetic_batch(C.make_cfg(), 2, 5), O.synthetic_batch(O.make_cfg(), 2, 5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_error_convention_without_gpu():
    """Argument validation happens before any CUDA call: bad arguments return a negative m2d_status and leave a
    message in m2d_last_error() (no exception crosses the C boundary, nothing is launched) — checkable on a CPU box."""
    import ctypes as C
    import __graft_entry__ as ge
    from music2dance_b200 import _lib
    ge.build()
    lib = _lib.load()
    a = _lib.RowConvArgs()                                  # all pointers NULL
    rc = lib.m2d_rowconv(C.byref(a), None)
    assert rc == -1 and b"null pointer" in lib.m2d_last_error()          # M2D_ERR_BAD_ARG
    w = _lib.WgradArgs()
    assert lib.m2d_wgrad(C.byref(w), None) == -1 and b"wgrad" in lib.m2d_last_error()
    assert lib.m2d_crop_batch(None, None, None, None, None, None, 1, 120, 69, 640, 76800, None, None, None) == -1
    assert lib.m2d_jerkiness(None, 1, 3, 69, None, None) == -1           # needs T > 3
    assert b"jerkiness" in lib.m2d_last_error()
    assert lib.m2d_adam(None, None, None, None, 0, None, 1e-3, 0.9, 0.999, 1e-8, 1.0, None) == -1
    with __import__("pytest").raises(RuntimeError, match="libm2d_b200"):
        _lib.call("m2d_gp_finalize_lp", None, 0, None, None, None)        # the Python wrapper raises with the message


def test_dropin_modules_refuse_cpu_tensors():
    """No CPU fallback: the drop-in entry points raise on CPU inputs instead of computing something else."""
    import pytest
    import torch
    from music2dance_b200.losses import gradient_penalty, jerkiness, tv_loss
    from music2dance_b200.utils import slice_audio_batch
    x = torch.zeros(2, 69, 120)
    for fn in (lambda: tv_loss(x), lambda: jerkiness(x),
               lambda: gradient_penalty(None, 2, x, x, x, is_seq=True, lp=False)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            slice_audio_batch(torch.zeros(2, 76800), 3200, 640, 2560)
    with pytest.raises(NotImplementedError):
        gradient_penalty(None, 2, x, x, is_seq=False)                      # phase1 branch lives in phase1.Phase1Trainer


def test_gemm_mode_names_match_header_enum():
    """ops.GEMM_MODES (the strings bench.py --gemm / M2D_GEMM accept) against the enum of include/m2d.h, and the
    mode setter's argument check (host-only: no kernel is launched)."""
    import re

    from music2dance_b200 import _lib, ops
    hdr = open(os.path.join(ROOT, "include", "m2d.h")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"M2D_GEMM_(\w+) = (\d+)", hdr))
    assert enum == {"FP32": 0, "TF32": 1, "TF32_BF16": 2, "TF32X3": 3}
    assert ops.GEMM_MODES == {"fp32": enum["FP32"], "tf32": enum["TF32"], "tf32bf16": enum["TF32_BF16"],
                              "tf32x3": enum["TF32X3"]}
    lib = _lib.load()
    before = lib.m2d_get_gemm_mode()
    try:
        for name, v in ops.GEMM_MODES.items():
            ops.set_gemm_mode(name)
            assert lib.m2d_get_gemm_mode() == v and ops.get_gemm_mode() == name
        assert lib.m2d_set_gemm_mode(7) != 0 and b"unknown mode" in lib.m2d_last_error()
    finally:
        lib.m2d_set_gemm_mode(before)
