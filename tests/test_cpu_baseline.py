"""The multithreaded CPU port (oracle/cpu_baseline.c, bit-vector Myers with cut-off and full trace
matrix) must agree bit for bit with the plain-DP oracle.  CPU only."""
import numpy as np
import pytest

from floxer_b200 import abi, synthetic
from floxer_b200.batch import VerifyConfig, alignment_records
from harness import oracle_align_tasks, oracle_verify_batch, random_align_tasks, results_as_tuples


@pytest.fixture(scope="module")
def baseline():
    from oracle import cpu_baseline
    cpu_baseline.lib()
    return cpu_baseline


@pytest.mark.parametrize("mode", [abi.MODE_EXISTS, abi.MODE_NO_CIGAR, abi.MODE_CIGAR])
@pytest.mark.parametrize("m_range,err_range", [((1, 70), (0.0, 0.3)), ((60, 200), (0.0, 0.15)), ((200, 700), (0.02, 0.12))])
def test_align_matches_oracle(oracle, baseline, mode, m_range, err_range):
    rng = np.random.default_rng(1000 + mode * 7 + m_range[0])
    ref, tasks, pool = random_align_tasks(rng, 150, m_range, err_range, mode, ref_len=20_000)
    res, cig = baseline.align_batch([ref], tasks, pool, threads=2)
    assert results_as_tuples(res, cig, tasks) == oracle_align_tasks(oracle, ref, tasks, pool)


@pytest.mark.parametrize("cfg", [
    VerifyConfig(interval_optimization=False),
    VerifyConfig(interval_optimization=True),
    VerifyConfig(interval_optimization=True, without_cigar=True),
    VerifyConfig(verification_kind=abi.KIND_DIRECT_FULL, interval_optimization=False),
    VerifyConfig(interval_optimization=True, extra_verification_ratio=0.3),
])
def test_verify_reads_matches_oracle(oracle, baseline, cfg):
    refs = [synthetic.random_reference(60_000, 11), synthetic.random_reference(30_000, 12)]
    batch = synthetic.make_batch(refs, 6, 700, 0.06, 99, oracle.pex_build, seed_errors=1, decoy_fraction=0.3)
    want, want_stats = oracle_verify_batch(oracle, refs, batch, cfg)
    al, cg, stats = baseline.verify_reads(refs, batch, cfg, threads=3)
    assert alignment_records(al, cg) == want
    assert stats == want_stats
    assert len(want) > 0
