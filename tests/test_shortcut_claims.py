"""The claims behind the host queue's shortcuts, checked on the CPU against the oracle:

* shared score passes: the minimum of row m of the UNION window's matrix over a member window's columns (and the rightmost
  column attaining it) is the member's own result whenever  end - m - score >= member start - union start,  and "no
  alignment" carries over unconditionally; the traceback from that cell of the union's matrix is the member's traceback;
* inner levels: an alignment found in the rightmost window A that ends inside a window B which starts at or before A lies
  in B; no alignment in the rightmost window A and none in the leftmost window C, A and C at most k + 1 apart, means none
  in any window in between.

Small alphabets and planted near-matches make ties and borderline cases frequent.  No GPU involved: these are properties
of the reference's alignment semantics (oracle.align = alignment::align), which is what makes the shortcuts exact."""
import numpy as np

from oracle import oracle as o

CIGAR_CONSUMES_REF = {"=": 1, "X": 1, "D": 1, "I": 0}


def last_row(window, query):
    """Row m of the semi-global matrix (free start and end in the window): value per column 0..n."""
    m = len(query)
    prev = np.arange(m + 1)
    out = [int(prev[m])]
    for c in window:
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = 0
        for i in range(1, m + 1):
            cur[i] = min(prev[i - 1] + (query[i - 1] != c), prev[i] + 1, cur[i - 1] + 1)
        out.append(int(cur[m]))
        prev = cur
    return out


def ref_span(cigar: str) -> int:
    n, num = 0, ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            n += int(num) * CIGAR_CONSUMES_REF[ch]
            num = ""
    return n


def planted_case(rng, alphabet):
    m = int(rng.integers(6, 28))
    k = int(rng.integers(0, max(1, m // 4) + 1))
    query = rng.integers(1, alphabet + 1, size=m, dtype=np.uint8)
    # a mutated copy of the query inside random text
    copy = list(query)
    for _ in range(int(rng.integers(0, k + 3))):
        kind = int(rng.integers(0, 3))
        pos = int(rng.integers(0, max(1, len(copy))))
        if kind == 0 and copy:
            copy[pos] = int(rng.integers(1, alphabet + 1))
        elif kind == 1:
            copy.insert(pos, int(rng.integers(1, alphabet + 1)))
        elif copy:
            del copy[pos]
    left = rng.integers(1, alphabet + 1, size=int(rng.integers(0, 2 * k + 8)), dtype=np.uint8)
    right = rng.integers(1, alphabet + 1, size=int(rng.integers(0, 2 * k + 8)), dtype=np.uint8)
    text = np.concatenate([left, np.array(copy, dtype=np.uint8), right]).astype(np.uint8)
    return query, text, k


def test_member_result_read_off_union_pass():
    rng = np.random.default_rng(20240601)
    decided = undecided = 0
    for _ in range(1500):
        query, union, k = planted_case(rng, int(rng.choice([2, 3, 4])))
        m, n_u = len(query), len(union)
        if n_u < 2:
            continue
        shift = int(rng.integers(0, min(n_u - 1, 2 * k + 4) + 1))
        n_b = int(rng.integers(1, n_u - shift + 1))
        member = union[shift:shift + n_b]
        own = o.align(member, query, k)
        row = last_row(union, query)
        cols = range(shift + 1, shift + n_b + 1)                      # the member's columns in the union's numbering
        L = min(row[c] for c in cols)
        e = max(c for c in cols if row[c] == L)
        if L > k:
            assert not own.exists                                    # none in the union's matrix there, none in the member's
            decided += 1
            continue
        if e - m - L >= shift:
            assert own.exists and own.num_errors == L
            assert own.start + ref_span(own.cigar) == e - shift      # the same (rightmost) end column ...
            # ... and the same traceback: the union window cut at column e has its rightmost minimum there or earlier;
            # ask the oracle for the alignment ending exactly at e by cutting the union at e and at the alignment's bound
            lo = e - m - L                                           # every optimal alignment ending at e starts at or after lo
            cut = o.align(union[lo:e], query, k)
            if cut.exists and cut.num_errors == L and cut.start + ref_span(cut.cigar) == e - lo:
                assert cut.cigar == own.cigar and lo + cut.start == shift + own.start
            decided += 1
        else:
            undecided += 1                                           # the queue scores such a member again on its own
    assert decided > 800 and undecided > 20


def test_inner_level_election_rules():
    rng = np.random.default_rng(20240602)
    yes_carried = no_carried = 0
    for _ in range(1500):
        query, text, k = planted_case(rng, int(rng.choice([2, 3])))
        m = len(query)
        base = m + 2 * k + 1                                         # window length of an inner node (verification.cpp:157-184)
        if len(text) < base + 2:
            text = np.concatenate([text, rng.integers(1, 3, size=base + 2 - len(text), dtype=np.uint8)])
        n = len(text)
        spread = int(rng.integers(0, k + 2))                         # A.start - C.start <= k + 1
        c_start = int(rng.integers(0, max(1, n - base - spread + 1)))
        a_start = min(c_start + spread, n - 1)

        def window(s):
            return text[s:min(s + base, n)]                          # clipped by the reference's end like the real ones
        a_res, c_res = o.align(window(a_start), query, k, mode=o.MODE_EXISTS), o.align(window(c_start), query, k, mode=o.MODE_EXISTS)
        for b_start in range(c_start, a_start + 1):
            b_res = o.align(window(b_start), query, k, mode=o.MODE_EXISTS)
            if not a_res.exists and not c_res.exists:
                assert not b_res.exists                              # none in A, none in C: none in between
                no_carried += 1
        full = o.align(window(a_start), query, k)                    # with positions, to know where A's alignment ends
        if full.exists:
            end = a_start + full.start + ref_span(full.cigar)
            for b_start in range(max(0, a_start - 2 * k - 3), a_start + 1):
                if min(b_start + base, n) >= end:
                    assert o.align(window(b_start), query, k, mode=o.MODE_EXISTS).exists
                    yes_carried += 1
        full_c = o.align(window(c_start), query, k)                  # level_third_kernel: an alignment found in C that provably lies in B
        if full_c.exists:
            end_c = c_start + full_c.start + ref_span(full_c.cigar)
            for b_start in range(c_start, a_start + 1):
                if end_c <= min(b_start + base, n) and end_c - m - full_c.num_errors >= b_start:
                    assert o.align(window(b_start), query, k, mode=o.MODE_EXISTS).exists
                    yes_carried += 1
    assert yes_carried > 500 and no_carried > 200
