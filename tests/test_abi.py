"""The C-ABI library builds for sm_100a, loads without a GPU and exports every
entry point include/nele_score.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nele_score.h")


@pytest.fixture(scope="module")
def lib_path():
    from nele_gan_b200 import build
    return build.build()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nele_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for need in ("nele_create", "nele_destroy", "nele_score_batch", "nele_last_error", "nele_get_stage",
                 "nele_last_timing", "nele_abi_version", "nele_set_profiling", "nele_kernel_time", "nele_features",
                 "nele_feature_frames"):
        assert need in names


def test_library_exports_every_declared_symbol(lib_path):
    dll = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(dll, name), name
    from nele_gan_b200 import engine
    assert sorted(engine.SYMBOLS) == declared_functions()
    assert dll.nele_abi_version() == 1


def test_constants_match_header():
    from nele_gan_b200 import engine
    src = open(HEADER).read()
    vals = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(NELE_[A-Z_0-9]+)\s+(-?0x[0-9a-fA-F]+|-?\d+)u?", src)}
    assert vals["NELE_METRIC_HASPI"] == engine.METRIC_HASPI
    assert vals["NELE_METRIC_SIIB"] == engine.METRIC_SIIB
    assert vals["NELE_METRIC_ESTOI"] == engine.METRIC_ESTOI
    assert vals["NELE_FLAG_MAPPED"] == engine.FLAG_MAPPED
    assert vals["NELE_FLAG_DEVICE_INPUT"] == engine.FLAG_DEVICE_INPUT
    assert vals["NELE_FLAG_NO_DITHER"] == engine.FLAG_NO_DITHER
    assert vals["NELE_FLAG_SIIB_NO_TILE"] == engine.FLAG_SIIB_NO_TILE
    assert vals["NELE_FLAG_KEEP_STAGES"] == engine.FLAG_KEEP_STAGES
    assert vals["NELE_FLAG_HASPI_V1"] == engine.FLAG_HASPI_V1
    assert vals["NELE_FLAG_STOI_CLASSIC"] == engine.FLAG_STOI_CLASSIC
    assert vals["NELE_FLAG_SIIB_KNN"] == engine.FLAG_SIIB_KNN
    assert vals["NELE_FLAG_HASQI_V2"] == engine.FLAG_HASQI_V2
    assert vals["NELE_ST_UNSUPPORTED"] == engine.ST_UNSUPPORTED
    assert vals["NELE_ST_TOO_SHORT"] == engine.ST_TOO_SHORT
    assert vals["NELE_INFO_SIIB_NULLSPACE"] == engine.INFO_SIIB_NULLSPACE


def test_no_gpu_fails_loudly(lib_path):
    """Without a CUDA device the engine refuses to exist: there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nele_gan_b200.engine import Engine, NeleError
    with pytest.raises(NeleError, match="no CUDA device|no CPU path|failed"):
        Engine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nele_gan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
