"""The ring rule of the DP engine on the shapes the host actually picks: a lane takes its next block only
after its own has ended, the band-limited work count agrees with the block geometry, and no ring is a lane longer than it
has to be.  Host arithmetic only -- no device."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def gpu():
    from floxer_b200 import build, gpu as g
    build.build_native()
    g.lib()
    return g


def block_ranges(n, m, k, W, nb):
    """Working columns [cs, ce] of every block (dp_kernels.cuh: set_block), None for a block outside the band."""
    R = 32 * W
    pad = nb * R - m
    dlo, dhi = -k - pad, n - m + k - pad
    out = []
    for b in range(nb):
        cs, ce = max(1, R * b + 1 + dlo), min(n, R * (b + 1) + dhi)
        out.append((cs, ce) if cs <= ce else None)
    return out


def shapes(rng, count):
    for _ in range(count):
        m = int(rng.integers(1, 30_000)) if rng.random() < 0.7 else int(rng.integers(1, 300))
        k = int(m * rng.uniform(0.0, 0.2))
        extra = int((m + 2 * k + 1) * rng.choice([0.0, 0.05, 0.3]))
        n = m + 2 * k + 1 + 2 * extra
        if rng.random() < 0.15:
            n = max(1, n - int(rng.integers(0, 2 * k + 2 * extra + 2)))      # a window clipped at the reference's end
        yield n, m, k


def test_ring_rule_holds_for_the_chosen_shapes(gpu):
    rng = np.random.default_rng(5)
    seen_rings = 0
    for count, (n, m, k) in enumerate(shapes(rng, 4000)):
        shape = gpu.engine_shape(n, m, k, with_traceback=bool(count & 1))
        if m - n > k:
            assert shape is None
            continue
        W, G, nb, ws = shape
        assert W in (1, 2, 4, 8, 16, 32) and nb == -(-m // (32 * W))
        ranges = block_ranges(n, m, k, W, nb)
        assert ws == sum((r[1] - r[0] + 1) * W for r in ranges if r)
        if G == 64:                                                          # the multi-warp kernel: a thread per block
            assert W == 32 and nb <= 128
            continue
        assert 1 <= G <= 32 and (nb == 1) == (G == 1)
        if G >= nb:
            continue                                                         # a lane per block
        seen_rings += 1
        for b in range(nb - G):
            if ranges[b] and ranges[b + G]:
                # block b's last step is ce + b; block b + G, on the same lane, begins with step cs + b + G
                assert ranges[b + G][0] + b + G > ranges[b][1] + b, (n, m, k, W, G, b)
        # every block begins while the block above is at work on the same column (the values left of its first column come
        # from there); with a band of a single diagonal the block above has ended on the column before
        for b in range(1, nb):
            if ranges[b]:
                up, cs = ranges[b - 1], ranges[b][0]
                assert up and up[0] <= cs, (n, m, k, W, b)
                assert cs <= up[1] or (n - m + 2 * k + 1 == 1 and cs - 1 == up[1]), (n, m, k, W, b)
    assert seen_rings > 1000


def test_rings_are_as_short_as_the_rule_allows(gpu):
    """G is 32 // (tasks per warp) for the largest number of tasks per warp whose ring still satisfies
    (32 W + 1) G >= 32 W + B - 1: one lane per ring fewer than in round 1 for most bands."""
    rng = np.random.default_rng(6)
    for n, m, k in shapes(rng, 2000):
        shape = gpu.engine_shape(n, m, k)
        if shape is None:
            continue
        W, G, nb, _ = shape
        if G == 64 or G >= nb:
            continue
        R, B = 32 * W, n - m + 2 * k + 1
        need = max(2, -(-(R + B - 1) // (R + 1))) if B > 4 else 3
        assert G >= min(need, nb)
        assert G == 32 // (32 // max(min(need, nb), 2)), (n, m, k, W, G, need)


def root_window(m, k, ratio=0.05):
    base = m + 2 * k + 1
    return base + 2 * int(np.ceil(base * ratio - 1e-9))


def test_traced_passes_count_their_traceback(gpu):
    """Config 2's root (5 kbp at 5 %): four tasks per warp at W = 8 beat two at W = 4 for the pass alone, but the traceback that
    follows every such pass costs more at W = 8 than the pass gains (profiles/r02_config2_issue_share_final.txt); the long
    roots of configs 3 and 4 keep their width either way."""
    n = root_window(5000, 250)
    assert gpu.engine_shape(n, 5000, 250)[:2] == (8, 8)
    assert gpu.engine_shape(n, 5000, 250, with_traceback=True)[:2] == (4, 16)
    for m, k in ((15000, 1200), (20000, 2000)):
        n = root_window(m, k)
        assert gpu.engine_shape(n, m, k)[0] == gpu.engine_shape(n, m, k, with_traceback=True)[0]
