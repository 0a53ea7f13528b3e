"""ctypes binding of the CPU oracle (oracle/floxer_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package floxer_b200 never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NULL_ID = 2**64 - 1

MODE_EXISTS, MODE_NO_CIGAR, MODE_CIGAR = 0, 1, 2
KIND_DIRECT_FULL, KIND_HIERARCHICAL = 0, 1
BUILD_RECURSIVE, BUILD_BOTTOM_UP = 0, 1
CIGAR_CHARS = {1: "I", 2: "D", 7: "=", 8: "X"}
CIGAR_CODES = {v: k for k, v in CIGAR_CHARS.items()}


class Node(C.Structure):
    _fields_ = [("parent_id", C.c_uint64), ("query_index_from", C.c_uint64),
                ("query_index_to", C.c_uint64), ("num_errors", C.c_uint64)]


class Anchor(C.Structure):
    _fields_ = [("pex_leaf_index", C.c_uint64), ("reference_id", C.c_uint64),
                ("reference_position", C.c_uint64), ("num_errors", C.c_uint64)]


class Span(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint64), ("extra", C.c_uint64)]


class Interval(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64)]


class Alignment(C.Structure):
    _fields_ = [("reference_id", C.c_uint64), ("start_in_reference", C.c_uint64),
                ("num_errors", C.c_uint64), ("orientation", C.c_uint32),
                ("cigar_len", C.c_uint32), ("cigar_offset", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_aligned_inner", "sum_aligned_inner", "n_aligned_root", "sum_aligned_root",
        "n_avoided_root", "sum_avoided_root", "cells_inner", "cells_root")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


NODE_DTYPE = np.dtype([("parent_id", "<u8"), ("query_index_from", "<u8"),
                       ("query_index_to", "<u8"), ("num_errors", "<u8")])
ANCHOR_DTYPE = np.dtype([("pex_leaf_index", "<u8"), ("reference_id", "<u8"),
                         ("reference_position", "<u8"), ("num_errors", "<u8")])


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and the CPU baseline) with gcc if missing or stale."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("floxer_oracle.c", "floxer_oracle.h", "cpu_baseline.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    outs = [so, os.path.join(_HERE, "libcpubaseline.so")]
    stale = force or any(not os.path.exists(o) for o in outs) or \
        max(os.path.getmtime(s) for s in srcs) > min(os.path.getmtime(o) for o in outs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    so = build()
    L = C.CDLL(so)
    u8p, u64p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.fxo_ceil_div.restype = C.c_uint64
    L.fxo_ceil_div.argtypes = [C.c_uint64, C.c_uint64]
    L.fxo_ceil_eps.restype = C.c_uint64
    L.fxo_ceil_eps.argtypes = [C.c_double]
    L.fxo_compute_span.restype = Span
    L.fxo_compute_span.argtypes = [C.c_uint64, C.POINTER(Node), C.c_uint64, C.c_uint64, C.c_double]
    L.fxo_interval_relationship.restype = C.c_int
    L.fxo_interval_relationship.argtypes = [Interval, Interval]
    L.fxo_interval_trim.restype = Interval
    L.fxo_interval_trim.argtypes = [Interval, C.c_uint64]
    L.fxo_intervals_new.restype = C.c_void_p
    L.fxo_intervals_new.argtypes = [C.c_int]
    L.fxo_intervals_free.argtypes = [C.c_void_p]
    L.fxo_intervals_configure.argtypes = [C.c_void_p, C.c_int]
    L.fxo_intervals_insert.argtypes = [C.c_void_p, Interval]
    L.fxo_intervals_contains.restype = C.c_int
    L.fxo_intervals_contains.argtypes = [C.c_void_p, Interval]
    L.fxo_intervals_size.restype = C.c_size_t
    L.fxo_intervals_size.argtypes = [C.c_void_p]
    L.fxo_pex_build.restype = C.c_int
    L.fxo_pex_build.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                C.POINTER(C.POINTER(Node)), C.POINTER(C.c_size_t),
                                C.POINTER(C.POINTER(Node)), C.POINTER(C.c_size_t)]
    L.fxo_free.argtypes = [C.c_void_p]
    L.fxo_align.restype = C.c_int
    L.fxo_align.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, C.c_int, u64p, u64p,
                            u32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.fxo_align_ex.restype = C.c_int
    L.fxo_align_ex.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_char_p, C.c_int,
                               u64p, u64p, u32p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.fxo_verifier_new.restype = C.c_void_p
    L.fxo_verifier_new.argtypes = [C.c_size_t, C.POINTER(u8p), u64p, C.POINTER(Node), C.c_size_t,
                                   C.POINTER(Node), C.c_size_t, C.c_int, C.c_int, C.c_double, C.c_int]
    L.fxo_verifier_free.argtypes = [C.c_void_p]
    L.fxo_verifier_run.restype = C.c_int
    L.fxo_verifier_run.argtypes = [C.c_void_p, u8p, C.c_size_t, C.c_int, C.POINTER(Anchor), C.c_size_t]
    L.fxo_verifier_configure_intervals.argtypes = [C.c_void_p, C.c_int]
    L.fxo_verifier_num_alignments.restype = C.c_size_t
    L.fxo_verifier_num_alignments.argtypes = [C.c_void_p]
    L.fxo_verifier_alignments.restype = C.POINTER(Alignment)
    L.fxo_verifier_alignments.argtypes = [C.c_void_p]
    L.fxo_verifier_cigar_pool.restype = u32p
    L.fxo_verifier_cigar_pool.argtypes = [C.c_void_p]
    L.fxo_verifier_stats.restype = C.POINTER(Stats)
    L.fxo_verifier_stats.argtypes = [C.c_void_p]
    _lib = L
    return L


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def _p8(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def cigar_to_string(ops) -> str:
    return "".join(f"{int(o) >> 4}{CIGAR_CHARS[int(o) & 15]}" for o in ops)


def cigar_from_string(s: str) -> list[int]:
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | CIGAR_CODES[ch])
            num = ""
    return out


@dataclass
class AlignResult:
    exists: bool
    num_errors: int = 0
    start: int = 0           # start in the window (add the span offset for start_in_reference)
    cigar: str = ""


def align(reference, query, max_errors: int, mode: int = MODE_CIGAR, priority: str | None = None,
          rightmost: bool = True) -> AlignResult:
    """alignment::align (src/lib/alignment.cpp:83-181) on rank sequences."""
    r, q = _u8(reference), _u8(query)
    n, m = len(r), len(q)
    cap = m + n + 2
    cig = np.zeros(cap, dtype=np.uint32)
    ne, st, cl = C.c_uint64(0), C.c_uint64(0), C.c_size_t(0)
    L = lib()
    args = [_p8(r), n, _p8(q), m, max_errors, mode]
    tail = [C.byref(ne), C.byref(st), cig.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(cl)]
    if priority is None and rightmost:
        rc = L.fxo_align(*args, *tail)
    else:
        rc = L.fxo_align_ex(*args, (priority or "LUD").encode(), int(rightmost), *tail)
    if rc < 0:
        raise RuntimeError("oracle align failed")
    if rc == 0:
        return AlignResult(False)
    return AlignResult(True, int(ne.value), int(st.value), cigar_to_string(cig[: cl.value]))


def compute_span(anchor_pos: int, node, leaf_from: int, ref_len: int, ratio: float):
    nd = Node(*[int(x) for x in node])
    s = lib().fxo_compute_span(anchor_pos, C.byref(nd), leaf_from, ref_len, ratio)
    return int(s.offset), int(s.length), int(s.extra)


def pex_build(total_len: int, num_errors: int, leaf_max_errors: int, strategy: int = BUILD_RECURSIVE):
    """Returns (inner, leaves) as structured numpy arrays with NODE_DTYPE (pex.cpp:84-256)."""
    pi, pl = C.POINTER(Node)(), C.POINTER(Node)()
    ni, nl = C.c_size_t(0), C.c_size_t(0)
    L = lib()
    if L.fxo_pex_build(total_len, num_errors, leaf_max_errors, strategy,
                       C.byref(pi), C.byref(ni), C.byref(pl), C.byref(nl)) != 0:
        raise RuntimeError("pex build failed")
    def grab(p, n):
        if n == 0:
            return np.zeros(0, dtype=NODE_DTYPE)
        buf = C.string_at(p, n * C.sizeof(Node))
        return np.frombuffer(buf, dtype=NODE_DTYPE).copy()
    inner, leaves = grab(pi, ni.value), grab(pl, nl.value)
    L.fxo_free(pi)
    L.fxo_free(pl)
    return inner, leaves


class Intervals:
    """intervals::verified_intervals (src/lib/intervals.cpp:78-127)."""

    def __init__(self, active: bool = True):
        self._h = lib().fxo_intervals_new(int(active))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fxo_intervals_free(self._h)
            self._h = None

    def configure(self, active: bool):
        lib().fxo_intervals_configure(self._h, int(active))

    def insert(self, start: int, end: int):
        lib().fxo_intervals_insert(self._h, Interval(start, end))

    def contains(self, start: int, end: int) -> bool:
        return bool(lib().fxo_intervals_contains(self._h, Interval(start, end)))

    def __len__(self):
        return int(lib().fxo_intervals_size(self._h))


def interval_relationship(a, b) -> int:
    return int(lib().fxo_interval_relationship(Interval(*a), Interval(*b)))


def interval_trim(a, amount: int):
    r = lib().fxo_interval_trim(Interval(*a), amount)
    return int(r.start), int(r.end)


class Verifier:
    """The per-read state of parallelization.cpp:120-129 plus query_verifier::verify per anchor."""

    def __init__(self, references, inner, leaves, kind=KIND_HIERARCHICAL, interval_optimization=True,
                 extra_verification_ratio=0.05, without_cigar=False):
        self._refs = [_u8(r) for r in references]
        n = len(self._refs)
        self._ptrs = (C.POINTER(C.c_uint8) * max(n, 1))(*[_p8(r) for r in self._refs])
        self._lens = (C.c_uint64 * max(n, 1))(*[len(r) for r in self._refs])
        self._inner = np.ascontiguousarray(inner, dtype=NODE_DTYPE)
        self._leaves = np.ascontiguousarray(leaves, dtype=NODE_DTYPE)
        self._h = lib().fxo_verifier_new(
            n, self._ptrs, self._lens,
            self._inner.ctypes.data_as(C.POINTER(Node)), len(self._inner),
            self._leaves.ctypes.data_as(C.POINTER(Node)), len(self._leaves),
            kind, int(interval_optimization), float(extra_verification_ratio), int(without_cigar))
        self._queries = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().fxo_verifier_free(self._h)
            self._h = None

    def configure_intervals(self, active: bool):
        lib().fxo_verifier_configure_intervals(self._h, int(active))

    def run(self, query, orientation: int, anchors):
        q = _u8(query)
        self._queries.append(q)
        a = np.ascontiguousarray(anchors, dtype=ANCHOR_DTYPE)
        rc = lib().fxo_verifier_run(self._h, _p8(q), len(q), orientation,
                                    a.ctypes.data_as(C.POINTER(Anchor)), len(a))
        if rc != 0:
            raise RuntimeError("oracle verify failed")

    def alignments(self):
        """[(reference_id, start_in_reference, num_errors, orientation, cigar_string)] in insertion order."""
        L = lib()
        n = L.fxo_verifier_num_alignments(self._h)
        al, pool = L.fxo_verifier_alignments(self._h), L.fxo_verifier_cigar_pool(self._h)
        out = []
        for i in range(n):
            a = al[i]
            ops = [pool[a.cigar_offset + k] for k in range(a.cigar_len)]
            out.append((int(a.reference_id), int(a.start_in_reference), int(a.num_errors),
                        int(a.orientation), cigar_to_string(ops)))
        return out

    def stats(self) -> dict:
        return lib().fxo_verifier_stats(self._h).contents.as_dict()
