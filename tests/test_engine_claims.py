"""The claim behind the DP engine's banding, checked on the CPU at the level of cells:

the matrix is cut into blocks of R rows; block b computes only the columns cs(b)..ce(b) that the band of diagonals
[-k, n - m + k] touches; everything it needs from outside is replaced by upper bounds -- its column cs - 1 is "the block
above's bottom value there, then +1 per row", the row above it reads "+1 per column" once the block above has ended (the
substitution the lane of the block BELOW makes, which is what lets a ring be one lane shorter) -- and still the minimum
of the last row inside the band, and the rightmost column attaining it, are the reference's result whenever an alignment
with at most k errors exists, and above k otherwise.  `model()` restates dp_kernels.cuh (set_block, the restart in
dp_task_group, run_steps' Upper) in plain integers; no GPU involved."""
import numpy as np

INF = 10 ** 9


def full_last_row(window, query):
    m = len(query)
    prev = list(range(m + 1))
    out = [prev[m]]
    for c in window:
        cur = [0] * (m + 1)
        for i in range(1, m + 1):
            cur[i] = min(prev[i - 1] + (query[i - 1] != c), prev[i] + 1, cur[i - 1] + 1)
        out.append(cur[m])
        prev = cur
    return out


def model(window, query, k, R):
    """(score, end column) as the engine computes them with blocks of R rows; score INF if the last row is never reached."""
    n, m = len(window), len(query)
    nb = -(-m // R)
    pad = nb * R - m                       # wildcard rows in front: row m is the last row of the last block
    dlo, dhi = -k - pad, n - m + k - pad
    q = [None] * pad + list(query)         # None matches everything
    bottom = {}                            # (block, column) -> value of the block's bottom row
    best, best_col = INF, 0
    for b in range(nb):
        cs, ce = max(1, R * b + 1 + dlo), min(n, R * (b + 1) + dhi)
        if cs > ce:
            continue
        rows = q[R * b: R * (b + 1)]
        if b == 0:
            col = [0 if r is None else None for r in rows]          # wildcard rows carry 0; below them +1 per row
            v = 0
            for i, r in enumerate(rows):
                v = 0 if r is None else v + 1
                col[i] = v
            top_prev = 0                                            # row 0 of a semi-global matrix
        else:
            # the block above is at work on column cs (its value there minus its last horizontal step = column cs - 1),
            # or -- a band of one diagonal -- it ended on column cs - 1
            if cs > 1:
                start = bottom[(b - 1, cs - 1)]                     # (always there: the block above begins no later and ends no earlier)
            else:
                start = max(0, R * b - pad)                         # column 0 of the matrix: value = row index (minus wildcards)
            col = [start + i + 1 for i in range(R)]
            top_prev = start
        for j in range(cs, ce + 1):
            if b == 0:
                top = 0
            elif (b - 1, j) in bottom:
                top = bottom[(b - 1, j)]
            else:
                top = top_prev + 1                                  # the block above has ended: "+1 per column"
            new = [0] * R
            for i, r in enumerate(rows):
                diag = (top_prev if i == 0 else col[i - 1]) + (0 if r is None or r == window[j - 1] else 1)
                up = (top if i == 0 else new[i - 1]) + 1
                new[i] = min(diag, up, col[i] + 1)
            col, top_prev = new, top
            bottom[(b, j)] = col[R - 1]
            if b == nb - 1 and col[R - 1] <= best:
                best, best_col = col[R - 1], j
    return best, best_col


def test_blocks_with_substituted_bounds_give_the_reference_result():
    rng = np.random.default_rng(11)
    checked = hits = 0
    for _ in range(1500):
        m = int(rng.integers(1, 70))
        k = int(rng.integers(0, max(1, m // 3) + 1))
        extra = int(rng.integers(0, 6))
        n = m + 2 * k + 1 + 2 * extra
        alphabet = int(rng.integers(2, 5))
        window = [int(x) for x in rng.integers(1, alphabet + 1, size=n)]
        if rng.random() < 0.7:                                      # a mutated copy of a piece of the window
            at = int(rng.integers(0, n - m + 1))
            query = window[at:at + m]
            for _ in range(int(rng.integers(0, k + 2))):
                kind, pos = int(rng.integers(0, 3)), int(rng.integers(0, max(1, len(query))))
                if kind == 0 and query:
                    query[pos] = int(rng.integers(1, alphabet + 1))
                elif kind == 1:
                    query.insert(pos, int(rng.integers(1, alphabet + 1)))
                elif len(query) > 1:
                    del query[pos]
        else:
            query = [int(x) for x in rng.integers(1, alphabet + 1, size=m)]
        m = len(query)
        if rng.random() < 0.15:                                     # a window clipped at the reference's end, or exactly the query's length
            n = max(1, int(rng.integers(max(1, m - k), n + 1)))
            window = window[:n]
        if m - n > k:
            continue
        row = full_last_row(window, query)
        lo, hi = max(0, m - k), min(n, n + k)                       # columns of the last row inside the band [-k, n - m + k]
        want = min(row[lo:hi + 1])
        for R in (4, 8, 16, 32):
            got, col = model(window, query, k, R)
            checked += 1
            if want <= k:
                hits += 1
                assert got == want, (window, query, k, R, got, want)
                assert col == max(j for j in range(lo, hi + 1) if row[j] == want), (window, query, k, R)
            else:
                assert got > k, (window, query, k, R, got, want)
    assert checked > 4000 and hits > 1500
