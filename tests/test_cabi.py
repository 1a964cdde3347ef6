"""CPU tier: the C-ABI library loads and exports every symbol include/marin_b200.h declares, and the
product fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "marin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    import prmers_b200 as pb
    L = pb.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(pb.exported_symbols()) == names
    assert b"sm_100a" in L.mb200_version()
    # the reference's own plugin ABI (third_party/aevum/src/EngineApi.h:28-62), exported by the same library
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "marin_b200.h")).read(), flags=re.S)
    aevum = sorted(set(re.findall(r"\b(aevum_engine_\w+)\s*\(", text)))
    assert len(aevum) == 19
    for n in aevum:
        assert hasattr(L, n), n


def test_planner_through_cabi_matches_oracle():
    import prmers_b200 as pb
    from oracle import marin_oracle as mo
    for p in (31, 127, 701, 86243, 1362763, 136279841, 175000001, 205271257, 600000001):
        assert pb.transform_size(p) == mo.transform_size(p)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import prmers_b200 as pb
    with pytest.raises(pb.MarinError, match="no CUDA device"):
        pb.Engine(127, 4)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "prmers_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f
