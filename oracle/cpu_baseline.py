"""ctypes binding of oracle/cpu_baseline.c -- the multithreaded CPU port used as the timed CPU baseline.

TEST / BENCH INFRASTRUCTURE ONLY (see the header of cpu_baseline.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from floxer_b200 import abi
from . import oracle as _oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _oracle.build()
        L = C.CDLL(os.path.join(_HERE, "libcpubaseline.so"))
        L.fxc_verify_reads.restype = C.c_int
        L.fxc_align_batch.restype = C.c_int
        L.fxc_result_num_alignments.restype = C.c_size_t
        L.fxc_result_alignments.restype = C.c_void_p
        L.fxc_result_cigar_len.restype = C.c_size_t
        L.fxc_result_cigar_pool.restype = C.c_void_p
        L.fxc_result_stats.restype = C.POINTER(abi.Stats)
        for f in (L.fxc_result_num_alignments, L.fxc_result_alignments, L.fxc_result_cigar_len,
                  L.fxc_result_cigar_pool, L.fxc_result_stats, L.fxc_result_free):
            f.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _refs(references):
    refs = [np.ascontiguousarray(r, dtype=np.uint8) for r in references]
    n = len(refs)
    ptrs = (C.c_void_p * max(n, 1))(*[r.ctypes.data for r in refs])
    lens = (C.c_uint64 * max(n, 1))(*[len(r) for r in refs])
    return refs, ptrs, lens


def verify_reads(references, batch, config, threads: int = 1):
    """Returns (alignments[abi.ALIGNMENT_DTYPE], cigar_pool[uint32], stats dict)."""
    refs, ptrs, lens = _refs(references)
    cfg = config.to_c()
    out = C.c_void_p()
    L = lib()
    rc = L.fxc_verify_reads(C.c_size_t(len(refs)), ptrs, lens, C.byref(cfg),
                            C.c_void_p(batch.reads.ctypes.data), C.c_size_t(len(batch.reads)),
                            C.c_void_p(batch.forward_pool.ctypes.data), C.c_void_p(batch.reverse_pool.ctypes.data),
                            C.c_void_p(batch.nodes.ctypes.data), C.c_void_p(batch.anchors.ctypes.data),
                            C.c_int(threads), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"cpu baseline failed: {rc}")
    try:
        n = L.fxc_result_num_alignments(out)
        al = np.frombuffer(C.string_at(L.fxc_result_alignments(out), n * abi.ALIGNMENT_DTYPE.itemsize),
                           dtype=abi.ALIGNMENT_DTYPE).copy() if n else np.zeros(0, dtype=abi.ALIGNMENT_DTYPE)
        nc = L.fxc_result_cigar_len(out)
        cg = np.frombuffer(C.string_at(L.fxc_result_cigar_pool(out), nc * 4), dtype=np.uint32).copy() if nc \
            else np.zeros(0, dtype=np.uint32)
        stats = L.fxc_result_stats(out).contents.as_dict()
    finally:
        L.fxc_result_free(out)
    return al, cg, stats


def align_batch(references, tasks, query_pool, inline_ref_pool=None, threads: int = 1, cigar_capacity=None):
    """Batched alignment::align on the CPU; returns (results[abi.ALIGN_RESULT_DTYPE], cigar_pool)."""
    refs, ptrs, lens = _refs(references)
    tasks = np.ascontiguousarray(tasks, dtype=abi.ALIGN_TASK_DTYPE)
    qp = np.ascontiguousarray(query_pool, dtype=np.uint8)
    ip = np.ascontiguousarray(inline_ref_pool if inline_ref_pool is not None else np.zeros(1, np.uint8), dtype=np.uint8)
    res = np.zeros(len(tasks), dtype=abi.ALIGN_RESULT_DTYPE)
    if cigar_capacity is None:
        cigar_capacity = int((tasks["ref_len"].astype(np.int64) + tasks["query_len"] + 2)[tasks["mode"] == 2].sum()) + 1
    cig = np.zeros(cigar_capacity, dtype=np.uint32)
    used = C.c_size_t(0)
    rc = lib().fxc_align_batch(C.c_size_t(len(refs)), ptrs, lens, C.c_void_p(tasks.ctypes.data), C.c_size_t(len(tasks)),
                               C.c_void_p(qp.ctypes.data), C.c_void_p(ip.ctypes.data), C.c_int(threads),
                               C.c_void_p(res.ctypes.data), C.c_void_p(cig.ctypes.data), C.c_size_t(cigar_capacity),
                               C.byref(used))
    if rc != 0:
        raise RuntimeError(f"cpu baseline align_batch failed: {rc}")
    return res, cig[: used.value]
