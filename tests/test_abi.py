"""The C-ABI boundary: the library builds/loads, exports every symbol include/letf_b200.h declares,
and the ctypes mirror (letf_monte_carlo_b200/abi.py) has the C struct layouts.  No device compute."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from letf_monte_carlo_b200 import abi
from letf_monte_carlo_b200.params import nearest_psd_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "letf_b200.h")


@pytest.fixture(scope="module")
def lib():
    from letf_monte_carlo_b200 import build
    build.build()                      # no-op when up to date; nvcc cross-compiles without a GPU
    return abi.load_engine_library()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(letf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_symbol_lists_agree():
    assert declared_functions() == sorted(abi.ENGINE_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.letf_abi_version() == abi.ABI_VERSION


def test_struct_layouts_match_the_header():
    names = {"LetfAssetParams": abi.AssetParams, "LetfStrategyParams": abi.StrategyParams,
             "LetfJointRegime": abi.JointRegime, "LetfVixRegime": abi.VixRegime,
             "LetfStressRegime": abi.StressRegime, "LetfFundingModel": abi.FundingModel,
             "LetfParams": abi.Params, "LetfTape": abi.Tape, "LetfOutputs": abi.Outputs,
             "LetfSummary": abi.Summary, "LetfLaunchInfo": abi.LaunchInfo, "LetfSeries": abi.Series}
    offs = [("LetfParams", "strategies"), ("LetfParams", "assets"), ("LetfParams", "joint"), ("LetfParams", "funding"),
            ("LetfParams", "duration_samples"), ("LetfTape", "te_t"), ("LetfOutputs", "jump_days"),
            ("LetfSummary", "bench_strategy"), ("LetfSeries", "regime")]
    prog = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void){']
    prog += [f'printf("{n} %zu\\n", sizeof({n}));' for n in names]
    prog += [f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));' for s, f in offs]
    prog += ['return 0;}']
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "sz.c"); exe = os.path.join(d, "sz")
        open(c, "w").write("\n".join(prog))
        subprocess.run(["gcc", "-std=c11", "-o", exe, c], check=True)
        got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for n, t in names.items():
        assert int(got[n]) == C.sizeof(t), n
    for s, f in offs:
        assert int(got[f"{s}.{f}"]) == getattr(names[s], f).offset, (s, f)


def test_nearest_psd3_matches_reference_formula(lib):
    """utils.py:42-53 (eigh, clip, renormalise) restated with a Jacobi solver for non-Python callers"""
    rng = np.random.default_rng(0)
    mats = [np.array([[1, .9, -.2], [.9, 1, -.18], [-.2, -.18, 1.0]]),
            np.array([[1, .99, -.9], [.99, 1, .9], [-.9, .9, 1.0]])]          # second one is indefinite
    for _ in range(20):
        a = rng.uniform(-1, 1, (3, 3)); a = (a + a.T) / 2; np.fill_diagonal(a, 1.0)
        mats.append(a)
    for m in mats:
        inp = np.ascontiguousarray(m, dtype=np.float64); out = np.zeros((3, 3))
        rc = lib.letf_nearest_psd3(inp.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        np.testing.assert_allclose(out, nearest_psd_matrix(m), atol=1e-12)


def test_invalid_arguments_fail_loudly(lib):
    import torch
    h = C.c_void_p()
    p = abi.Params()                                     # abi_version 0 -> rejected before any CUDA call
    rc = lib.letf_create(C.byref(p), 0, abi.PREC_F32, C.byref(h))
    assert rc == abi.ERR_INVALID and b"abi_version" in lib.letf_last_error()
    assert lib.letf_run(None, None, 0, 0, 1, 0, None, None) == abi.ERR_INVALID
    if not torch.cuda.is_available():
        # no device: the product refuses to run instead of falling back to a CPU path
        from letf_monte_carlo_b200 import synthetic
        from letf_monte_carlo_b200.engine import PathEngine
        from letf_monte_carlo_b200.params import build_params
        params, keep = build_params(["S1"], 1, synthetic.fallback_regime_model(), synthetic.config_namespace())
        rc = lib.letf_create(C.byref(params), 0, abi.PREC_F32, C.byref(h))
        assert rc == abi.ERR_CUDA and lib.letf_last_error()
        with pytest.raises(RuntimeError):
            PathEngine(params, keep)


def test_missing_library_raises(tmp_path):
    with pytest.raises(abi.EngineLibraryMissing):
        abi.load_engine_library(str(tmp_path / "libletf_b200.so"))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "letf_monte_carlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "letf_oracle" not in txt, f
