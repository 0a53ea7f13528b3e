"""Synthetic workloads of the shapes BASELINE.json names.

Reads follow the reference simulator's model (src/main/simulated_dataset.cpp:81-223): exactly
floor(rate * base_length) edits at distinct origin positions, kind uniform over {mismatch, insertion after,
deletion}, mismatch base != origin base, inserted base uniform.

Anchors are a labelled STAND-IN for search::searcher::search_seeds (src/lib/search.cpp:143-324), which
cannot be built offline: every leaf whose span carries <= leaf.num_errors edits gets one anchor at the
origin-mapped reference position; a configurable fraction of leaves additionally gets a uniformly random
decoy anchor (the false positives the real seeder reports); order is seed -> reference -> position
(src/lib/search.cpp:78-100).
"""
from __future__ import annotations

import math

import numpy as np

from . import abi
from .batch import BatchBuilder, ReadBatch

COMPLEMENT = np.array([0, 4, 3, 2, 1, 5], dtype=np.uint8)


def ceil_eps(value: float) -> int:
    """math::floating_point_error_aware_ceil (include/math.hpp:22-27)."""
    eps = 0.000000001
    return int(math.ceil(value - eps) + eps)


def reverse_complement(r: np.ndarray) -> np.ndarray:
    return COMPLEMENT[r][::-1].copy()


def random_reference(length: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(1, 5, size=length, dtype=np.uint8)


def plant_repeats(ref: np.ndarray, seed: int, families: int, unit=(500, 5000), copies=(5, 50), divergence=0.02):
    """Config-3 style repeat families: overwrite random loci with mutated copies of a unit."""
    rng = np.random.default_rng(seed)
    n = len(ref)
    for _ in range(families):
        ulen = int(rng.integers(unit[0], unit[1] + 1))
        unit_seq = rng.integers(1, 5, size=ulen, dtype=np.uint8)
        for _ in range(int(rng.integers(copies[0], copies[1] + 1))):
            cp = unit_seq.copy()
            nmut = int(divergence * ulen)
            pos = rng.choice(ulen, size=nmut, replace=False)
            cp[pos] = ((cp[pos] - 1 + rng.integers(1, 4, size=nmut)) % 4 + 1).astype(np.uint8)
            at = int(rng.integers(0, n - ulen))
            ref[at:at + ulen] = cp
    return ref


def simulate_read(rng, ref: np.ndarray, start: int, base_len: int, n_err: int):
    """Returns (read ranks in reference orientation, origin position per read base, dirty flag per read base).

    A base is dirty if it is a mismatch, an inserted base, or the base that follows a deletion."""
    frag = ref[start:start + base_len]
    idx = np.sort(rng.choice(base_len, size=n_err, replace=False))
    kinds = rng.integers(0, 3, size=n_err)                # 0 mismatch, 1 insertion after, 2 deletion
    mi, ii, di = idx[kinds == 0], idx[kinds == 1], idx[kinds == 2]
    emit = np.ones(base_len, dtype=np.int64)
    emit[di] = 0
    emit[ii] = 2
    offs = np.cumsum(emit) - emit
    total = int(emit.sum())
    out = np.empty(total, dtype=np.uint8)
    origin = np.empty(total, dtype=np.int64)
    dirty = np.zeros(total, dtype=bool)
    keep = emit > 0
    pos = np.arange(base_len, dtype=np.int64)
    out[offs[keep]] = frag[keep]
    origin[offs[keep]] = start + pos[keep]
    nb = rng.integers(0, 3, size=len(mi))
    nb = nb + (nb >= frag[mi].astype(np.int64) - 1)       # choose_distinct_rank, simulated_dataset.cpp:75-79
    out[offs[mi]] = (nb + 1).astype(np.uint8)
    dirty[offs[mi]] = True
    out[offs[ii] + 1] = rng.integers(1, 5, size=len(ii), dtype=np.uint8)
    origin[offs[ii] + 1] = start + ii + 1
    dirty[offs[ii] + 1] = True
    if total:
        dirty[np.minimum(offs[di], total - 1)] = True
    return out, origin, dirty


def make_batch(references, n_reads: int, base_len: int, error_rate: float, seed: int, pex_build,
               seed_errors: int = 2, decoy_fraction: float = 0.25, bottom_up: bool = False,
               alternate_strands: bool = True, max_anchors_per_leaf: int = 50) -> ReadBatch:
    """Simulated reads + PEX trees + stand-in anchors for `references` (list of uint8 rank arrays).

    pex_build(total_len, num_errors, leaf_max_errors, strategy) -> (inner, leaves) structured arrays."""
    rng = np.random.default_rng(seed)
    n_err = int(error_rate * base_len)
    bb = BatchBuilder()
    ref_lens = [len(r) for r in references]
    truth = []
    for ri in range(n_reads):
        rid = int(rng.integers(0, len(references)))
        ref = references[rid]
        start = int(rng.integers(0, ref_lens[rid] - base_len - 1))
        seq, origin, dirty = simulate_read(rng, ref, start, base_len, n_err)
        on_reverse = alternate_strands and (ri % 2 == 1)
        # `seq` is in reference orientation; a read sequenced from the reverse strand is its reverse complement
        fwd = reverse_complement(seq) if on_reverse else seq
        rc = seq if on_reverse else reverse_complement(seq)
        k = ceil_eps(len(fwd) * error_rate)                      # input.cpp:26-34
        inner, leaves = pex_build(len(fwd), k, seed_errors, 1 if bottom_up else 0)
        csum = np.concatenate([[0], np.cumsum(dirty)])
        true_anchors, decoys = [], []
        for li, leaf in enumerate(leaves):
            f, t, e = int(leaf["query_index_from"]), int(leaf["query_index_to"]), int(leaf["num_errors"])
            d = int(csum[t + 1] - csum[f])
            if d <= e:
                true_anchors.append((li, rid, int(origin[f]), d))
            if rng.random() < decoy_fraction:
                drid = int(rng.integers(0, len(references)))
                decoys.append((li, drid, int(rng.integers(0, ref_lens[drid] - (t - f + 1))), e))
        same = sorted(true_anchors + decoys[0::2])[: max_anchors_per_leaf * len(leaves)]
        other = sorted(decoys[1::2])
        af, ar = (other, same) if on_reverse else (same, other)
        bb.add(fwd, rc, inner, leaves, np.array(af, dtype=abi.ANCHOR_DTYPE), np.array(ar, dtype=abi.ANCHOR_DTYPE))
        truth.append((rid, start, on_reverse))
    return bb.build(truth=truth, base_len=base_len, error_rate=error_rate, seed=seed)


def microbench_tasks(reference: np.ndarray, lengths, error_rates, tasks_per_cell: int, seed: int, mode: int,
                     positive_fraction: float = 0.5):
    """Config 5: batched edit distance, query m x window n = m + 2k + 1.

    Returns (tasks[abi.ALIGN_TASK_DTYPE], query_pool)."""
    rng = np.random.default_rng(seed)
    tasks, pool, off = [], [], 0
    n_ref = len(reference)
    for m in lengths:
        for e in error_rates:
            k = math.ceil(m * e)
            n = m + 2 * k + 1
            for _ in range(tasks_per_cell):
                at = int(rng.integers(0, n_ref - n))
                if rng.random() < positive_fraction:
                    q, _, _ = simulate_read(rng, reference, at + k, m, int(m * e))
                else:
                    q = rng.integers(1, 5, size=m, dtype=np.uint8)
                pool.append(q)
                tasks.append((at, at, off, n, len(q), 0, k, mode, 0, (0,) * 6))
                off += len(q)
    return np.array(tasks, dtype=abi.ALIGN_TASK_DTYPE), np.concatenate(pool)
