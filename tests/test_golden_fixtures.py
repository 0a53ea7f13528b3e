"""tests/golden/oracle_cases.json: the oracle still reproduces it (CPU), the CUDA path produces it (GPU)."""
import importlib.util
import json
import os

import pytest

from floxer_b200.batch import VerifyConfig, alignment_records
from harness import results_as_tuples

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(HERE, "golden", "make_fixtures.py"))
MF = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MF)


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "oracle_cases.json")) as f:
        return json.load(f)


def _norm(x):
    return json.loads(json.dumps(x))


def test_oracle_reproduces_the_fixture(oracle, golden):
    assert _norm(MF.build()) == golden


@pytest.mark.gpu
def test_cuda_path_reproduces_the_fixture(golden):
    from floxer_b200 import build, gpu as g
    build.build_native()
    ctx = g.Context(0)
    try:
        refs, batch = MF.verify_inputs()
        ctx.set_references(refs)
        for case in MF.VERIFY_CASES:
            job = ctx.verify_reads(batch, VerifyConfig(**case["cfg"]))
            al, cg = job.alignments()
            want = golden["verify"][case["name"]]
            assert _norm([list(r) for r in alignment_records(al, cg)]) == want["records"], case["name"]
            assert job.stats() == want["stats"], case["name"]
            job.free()
        for case in MF.ALIGN_CASES:
            ref, tasks, pool = MF.align_inputs(case)
            ctx.set_references([ref])
            res, cig = ctx.align_batch(tasks, pool)
            assert _norm([list(r) for r in results_as_tuples(res, cig, tasks)]) == golden["align"][case["name"]], case["name"]
    finally:
        ctx.close()
