"""CPU-only checks of the product library: it loads, exports every symbol include/floxer_gpu.h declares,
its host-side PEX builder agrees with the oracle and the reference's golden trees, and it fails loudly
(not silently on a CPU path) when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_vectors as G
from floxer_b200 import abi, build, gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_native()
    return gpu.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "floxer_gpu.h")).read()
    declared = set(re.findall(r"\b(fxg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/floxer_gpu.h but not exported"
    assert declared == set(gpu.EXPORTS)


def test_struct_layouts_match_header():
    assert abi.PEX_NODE_DTYPE.itemsize == 32 and abi.ANCHOR_DTYPE.itemsize == 32
    assert abi.ALIGN_TASK_DTYPE.itemsize == 48 and abi.ALIGN_RESULT_DTYPE.itemsize == 32
    assert abi.READ_DTYPE.itemsize == 48 and abi.ALIGNMENT_DTYPE.itemsize == 40
    assert C.sizeof(abi.VerifyConfig) == 16 and C.sizeof(abi.Stats) == 64 and C.sizeof(abi.Counters) == 184


def test_pex_builder_golden(lib):
    strat = {"recursive": 0, "bottom_up": 1}
    for (total, errs, leaf_errs, s), want in G.PEX_LEAVES:
        inner, leaves = gpu.pex_build(total, errs, leaf_errs, strat[s])
        got = [(int(l["query_index_from"]), int(l["query_index_to"] - l["query_index_from"] + 1), int(l["num_errors"]))
               for l in leaves]
        assert got == want


@pytest.mark.parametrize("strategy", [0, 1])
def test_pex_builder_matches_oracle(lib, oracle, strategy):
    rng = np.random.default_rng(5 + strategy)
    cases = [(12, 3, 0), (30, 5, 1), (5000, 250, 2), (15000, 1200, 2), (20000, 2000, 2), (100000, 10000, 3), (64, 1, 2)]
    cases += [(int(n), int(max(1, n * e)), int(s)) for n, e, s in
              zip(rng.integers(20, 30000, 40), rng.uniform(0.01, 0.2, 40), rng.integers(0, 4, 40))]
    for total, errs, leaf_errs in cases:
        if errs >= total:
            continue
        a = gpu.pex_build(total, errs, leaf_errs, strategy)
        b = oracle.pex_build(total, errs, leaf_errs, strategy)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (total, errs, leaf_errs)


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created; nothing computes on the host instead."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.fxg_create(0, C.byref(h)) != 0 and not h.value
    with pytest.raises(gpu.FloxerGpuError):
        gpu.Context(0)
    assert lib.fxg_pex_build(0, 0, 0, 0, None, None, None, None) == abi.ERR_INVALID_ARGUMENT
