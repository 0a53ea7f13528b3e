#!/usr/bin/env python
"""bench.py -- PEX hierarchical verification throughput (BASELINE.json metric) on N B200s of one node.

One "step" = one pass of the hot path over one batch of synthetic reads (config 2 of BASELINE.json:
10 Mbp random reference, 1 000 simulated 5 kbp reads at 5 % error, floxer defaults) per GPU.  Reads shard
across ranks with no data-path collective (weak scaling: every rank verifies its own 1 000-read batch).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA path through the C ABI)
  python bench.py --impl reference [...]                         CPU arm: the multithreaded CPU port of the
                                                                 reference path (the reference binary cannot be
                                                                 built offline) on a bounded sample

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the library asks for 32 hardware work queues when it is loaded (floxer_gpu.cu, fxg_on_load); torch may create the CUDA
# context before that, so the same default is set here, before torch is imported
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (reference length, reads per GPU, read length, error rate, reference seed, read seed)
    "config2": dict(ref_len=10_000_000, reads=1000, read_len=5000, error=0.05, ref_seed=20240001, read_seed=20240003),
    "config2_small": dict(ref_len=2_000_000, reads=64, read_len=5000, error=0.05, ref_seed=20240001, read_seed=20240003),
}
MYERS_INSTR_PER_WORD_STEP = 11          # minimal LOP3/IADD3/SHF sequence of one 32-cell word-step


def make_workload(name: str, rank: int, pex_build):
    from floxer_b200 import synthetic
    w = WORKLOADS[name]
    refs = [synthetic.random_reference(w["ref_len"], w["ref_seed"])]
    batch = synthetic.make_batch(refs, w["reads"], w["read_len"], w["error"], w["read_seed"] + 1000 * rank, pex_build,
                                 seed_errors=2, decoy_fraction=0.25)
    return refs, batch


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def window(self, t0: float, t1: float):
        """A timed region (time.perf_counter values): only samples taken inside one count, if any were."""
        self.windows.append((t0, t1))

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        inside = [r for t, r in self.rows if any(a <= t <= b + 0.1 for a, b in self.windows)]
        rows = inside if inside else [r for _, r in self.rows]      # a region shorter than the sampling period: the warm-up's
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_sample_size(refs, batch, cfg, threads: int, target_s: float = 12.0) -> int:
    """Sizes the CPU sample to about `target_s` seconds of wall time from a short probe."""
    from oracle import cpu_baseline
    probe = batch.slice(0, min(len(batch), max(4, threads // 2)))
    t0 = time.perf_counter()
    cpu_baseline.verify_reads(refs, probe, cfg, threads=threads)
    rate = len(probe) / max(time.perf_counter() - t0, 1e-6)
    return int(max(min(len(batch), rate * target_s), min(len(batch), threads)))


def cpu_arm(refs, batch, cfg, sample_reads: int, threads: int, repeats: int = 1):
    """Times the multithreaded CPU port on the first `sample_reads` reads; returns (seconds, stats)."""
    from oracle import cpu_baseline
    if not sample_reads:
        sample_reads = cpu_sample_size(refs, batch, cfg, threads)
    sample = batch.slice(0, min(sample_reads, len(batch)))
    best, stats = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, _, stats = cpu_baseline.verify_reads(refs, sample, cfg, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, stats, len(sample)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=192)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--interval-optimization", action="store_true", help="floxer --interval-optimization (off by default, as in the reference)")
    ap.add_argument("--cpu-sample-reads", type=int, default=0, help="reads in the CPU sample (0 = sized automatically)")
    ap.add_argument("--pipeline", type=int, default=32, help="batches in flight per GPU: the library serves FXG_GROUPS (default and at most 32) *_run calls at a time")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    threads = os.cpu_count() or 1

    from floxer_b200.batch import VerifyConfig
    cfg = VerifyConfig(interval_optimization=args.interval_optimization)
    W = WORKLOADS[args.workload]
    config = {"workload": f"{args.workload}: {W['ref_len']} bp uniform random reference, {W['reads']} simulated reads x "
                          f"{W['read_len']} bp at {int(W['error'] * 100)} % error per GPU, recursive PEX tree, seed errors 2, "
                          f"hierarchical verification, interval optimization {'on' if cfg.interval_optimization else 'off'}, "
                          f"extra verification ratio 0.05, CIGAR output; anchors from the ground-truth stand-in seeder",
              "reads_per_gpu": W["reads"], "read_len": W["read_len"], "error_rate": W["error"],
              "l2": "256 MiB written to HBM between steps (twice the 126 MB L2); the batches in flight run concurrently, so a step never finds its own data in L2",
              "batches_in_flight": max(1, min(args.pipeline, int(os.environ.get("FXG_GROUPS", "32"))))}

    # ------------------------------------------------------------------ CPU arm ("reference")
    if args.impl == "reference":
        if rank != 0:
            return 0
        from floxer_b200 import build
        from floxer_b200 import gpu as g          # host-side PEX builder only (no device needed)
        build.build_native()
        refs, batch = make_workload(args.workload, 0, g.pex_build)
        # every step is a sample of the workload sized so that the K steps together take about three minutes
        n_sample = args.cpu_sample_reads or cpu_sample_size(refs, batch, cfg, threads, target_s=min(10.0, max(0.3, 200.0 / max(args.steps, 1))))
        times, stats = [], None
        for _ in range(max(args.steps, 1)):
            dt, stats, n_used = cpu_arm(refs, batch, cfg, n_sample, threads)
            times.append(dt)
        sec = sum(times) / len(times)
        cells = stats["cells_inner"] + stats["cells_root"]
        gcups = cells / sec / 1e9
        line = {"impl": "reference", "metric": "pex_verification_gcups", "value": gcups, "unit": "GCUPS",
                "reads_per_s": n_used / sec, "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port",
                                 "sample": f"first {n_used} reads of the workload per step, all anchors, both strands, CIGARs",
                                 "reads_per_s": n_used / sec},
                "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "CPU port of floxer 0.2.0 + SeqAn3 edit-distance path (oracle/cpu_baseline.c); the reference binary "
                        "cannot be built offline"}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the product path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when it initialises: keep stdout for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from floxer_b200 import build
    from floxer_b200 import gpu as g
    build.build_native()
    refs, batch = make_workload(args.workload, rank, g.pex_build)
    # the step's inputs live in page-locked host memory (the host->device copies inside the timed e2e region are plain DMA)
    pinned = [torch.from_numpy(a).pin_memory() for a in (batch.forward_pool, batch.reverse_pool)]
    batch.forward_pool, batch.reverse_pool = pinned[0].numpy(), pinned[1].numpy()
    ctx = g.Context(local_rank)
    ctx.set_references(refs)
    int32_peak = ctx.measure_int32_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    depth = max(1, min(args.pipeline, int(os.environ.get("FXG_GROUPS", "32"))))

    def run_lanes(step_fns, n_steps):
        """n_steps steps, dealt round-robin to len(step_fns) host threads (batches in flight); returns wall seconds."""
        counts = [n_steps // len(step_fns) + (1 if i < n_steps % len(step_fns) else 0) for i in range(len(step_fns))]
        errors = []

        def lane(i):
            try:
                for _ in range(counts[i]):
                    step_fns[i]()
            except Exception as e:                       # noqa: BLE001 -- re-raised by the main thread
                errors.append(e)
        threads = [threading.Thread(target=lane, args=(i,)) for i in range(len(step_fns))]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        return dt

    # ---- device-resident arm: inputs staged once (one staged copy per batch in flight), each step = fxg_verify_run ----
    sampler = ClockSampler(local_rank)
    if rank == 0 and os.environ.get("BENCH_NO_SAMPLER") != "1":      # rank 0 reports its GPU's clocks (one nvidia-smi, not one per rank)
        sampler.start()
    jobs = [ctx.stage_verify(batch, cfg) for _ in range(depth)]
    run_lanes([j.run for j in jobs], args.warmup * depth)    # both worker groups warm (their buffers are allocated on first use)

    def resident_step(j):
        def f():
            flush.fill_(1)                               # 256 MiB written between steps (queued on torch's stream, no host sync)
            j.run()
        return f
    # one step at a time first: the latency of a step, and its device time between the run's first and last operation
    lat_ms, kernel_ms = [], []
    for it in range(3):
        flush.fill_(it & 0xff)
        torch.cuda.synchronize()
        c0 = ctx.counters()
        t0 = time.perf_counter()
        jobs[0].run()
        lat_ms.append((time.perf_counter() - t0) * 1e3)
        kernel_ms.append(ctx.counters()["run_ms"] - c0["run_ms"])
    barrier()
    ctx.reset_counters()
    t_region = time.perf_counter()
    total_s = run_lanes([resident_step(j) for j in jobs], args.steps)
    barrier()
    sampler.window(t_region, time.perf_counter())
    step_ms = [total_s * 1e3 / args.steps] * args.steps
    ctr = ctx.counters()
    stats = jobs[0].stats()
    al, cg = jobs[0].alignments()
    n_alignments = len(al)
    for j in jobs:
        j.free()

    # ---- end-to-end arm: host buffers in, alignments + CIGARs out, every step ----
    checksum = [0] * depth

    def e2e_step(i):
        def f():
            flush.fill_(2)
            j2 = ctx.verify_reads(batch, cfg)            # host buffers in: H2D, all waves, tracebacks, D2H of alignments + CIGARs
            a2, c2 = j2.alignments(copy=False)           # what a C caller reads: the job's own result arrays
            checksum[i] ^= int(a2["start_in_reference"].sum()) ^ int(a2["num_errors"].sum()) ^ len(c2)
            del a2, c2
            j2.free()
        return f
    run_lanes([e2e_step(i) for i in range(depth)], 2 * depth)   # warm-up (page-locked pools are allocated once)
    torch.cuda.synchronize()
    ctx.reset_counters()
    n_e2e = args.steps
    t_region = time.perf_counter()
    e2e_s = run_lanes([e2e_step(i) for i in range(depth)], n_e2e)
    torch.cuda.synchronize()
    sampler.window(t_region, time.perf_counter())
    clocks = sampler.stop()
    e2e_ms = [e2e_s * 1e3 / n_e2e] * n_e2e
    e2e_ctr = ctx.counters()

    # ---- roofline passes (rank 0): one host worker / one stream, so that kernels do not overlap and the
    #      CUDA-event time of the DP launches is the time of those launches alone ----
    n_roof = 3

    def roofline_pass(extra_env):
        env = {"FXG_WORKERS": "1", "FXG_GROUPS": "1", **extra_env}
        saved_env = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        ctx1 = g.Context(local_rank)
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        ctx1.set_references(refs)
        job1 = ctx1.stage_verify(batch, cfg)
        job1.run()
        ctx1.reset_counters()
        for it in range(n_roof):
            flush.fill_(it & 0xff)
            torch.cuda.synchronize()
            job1.run()
        out = ctx1.counters()
        job1.free()
        ctx1.close()
        return out

    roof_ctr = prod_ctr = None
    if rank == 0:
        # the engine on a launch that fills the machine: every root window scored on its own, as when the windows of a
        # batch do not coincide (14 542 score passes in one launch) ...
        roof_ctr = roofline_pass({"FXG_SHARE_ROOTS": "0", "FXG_INFER_INNER": "0"})
        # ... and the launches of the step as benchmarked (coinciding windows share one pass: 7x fewer word-steps)
        prod_ctr = roofline_pass({})

    # max over ranks
    my = torch.tensor([sum(step_ms), sum(e2e_ms) / n_e2e * args.steps, sum(kernel_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(my, op=dist.ReduceOp.MAX)
    total_ms, e2e_total_ms, dev_ms = [float(x) for x in my.cpu()]
    cells_step = stats["cells_inner"] + stats["cells_root"]
    units = torch.tensor([cells_step, len(batch)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    cells_all, reads_all = [float(x) for x in units.cpu()]

    ms_per_step = total_ms / args.steps
    gcups = cells_all / (ms_per_step * 1e-3) / 1e9
    e2e_ms_per_step = e2e_total_ms / args.steps
    e2e_gcups = cells_all / (e2e_ms_per_step * 1e-3) / 1e9

    line = None
    if rank == 0:
        # roofline of the dominant kernel (the bit-vector DP engine): integer-ALU bound, SURVEY 8(d).
        # `frac` is that of the engine's root-level dp_kernel launch with every root window of the batch scored on its own
        # (a launch that fills the machine), timed alone with its own CUDA event pair on its stream; `all_launches` is
        # every score-pass launch of that step over the event time of the waves (small launches and their tails included);
        # `as_benchmarked` is the same pair of figures for the step the headline numbers time, where coinciding windows
        # share one pass: one batch alone then leaves most of the machine idle (the launches are chains of dependent
        # steps, a few hundred warps wide), and the batches in flight fill it together.
        dp_s = roof_ctr["dp_kernel_ms"] * 1e-3
        ws = roof_ctr["dp_word_steps"]
        root_s = roof_ctr["root_launch_ms"] * 1e-3
        root_ws = roof_ctr["root_launch_word_steps"]
        achieved = root_ws * MYERS_INSTR_PER_WORD_STEP / root_s if root_s > 0 else 0.0
        achieved_all = ws * MYERS_INSTR_PER_WORD_STEP / dp_s if dp_s > 0 else 0.0
        tr_s = roof_ctr["trace_kernel_ms"] * 1e-3
        roofline = {"bound": "int32_alu", "achieved": achieved / 1e9, "peak": int32_peak / 1e9, "unit": "Ginstr/s",
                    "frac": achieved / int32_peak if int32_peak else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this launch, one `ncu --set full` capture
                    # (profiles/r01_end_unshared_ncu_full_summary.txt: 84.4 MB read, 1.446 GB written -- the checkpoint records)
                    "traffic": 1530833704,
                    # the contract's HBM view of the same launch, for the record: the path is not bandwidth-bound
                    "hbm_view": (lambda peak, src: {"achieved": 1530833704 / root_s * n_roof / 1e9 if root_s > 0 else None, "peak": peak, "unit": "GB/s",
                                                    "frac": (1530833704 / root_s * n_roof / 1e9) / peak if root_s > 0 else None,
                                                    "peak_source": src,
                                                    "note": "DRAM bytes of the launch (ncu) / its CUDA-event time; 5-6 % of the copy bandwidth"})(
                        *((_measured_peak("hbm_gbs"), "MEASURED_PEAKS.json hbm_gbs") if _measured_peak("hbm_gbs") else (6650.0, "of fallback (B200_PROFILING.md)"))),
                    "kernel": "fxg::dp_kernel<4,true> -- the root-level launch of a step (score pass leaving traceback checkpoints), "
                              "every root window of the batch scored on its own (FXG_SHARE_ROOTS=0 FXG_INFER_INNER=0): the launch that fills the machine",
                    "how": "algorithmic 11 int32 instr per 32-cell word-step x word-steps issued by that launch (band-limited, counted on the "
                           "host from the band geometry) / CUDA-event time of the launch on its stream; peak = LOP3/IADD3/SHF 8:1:2 "
                           "issue-rate microbenchmark on this GPU in this run; HBM traffic is negligible (window + Eq words in, "
                           "one checkpoint record per block and 32 steps out)",
                    "word_steps_per_launch": root_ws / n_roof, "launch_ms": roof_ctr["root_launch_ms"] / n_roof,
                    "all_launches": {"frac": achieved_all / int32_peak if int32_peak else None, "achieved": achieved_all / 1e9,
                                     "dp_word_steps_per_step": ws / n_roof, "dp_kernel_ms_per_step": roof_ctr["dp_kernel_ms"] / n_roof},
                    "cells_computed_per_step": ws * 32 / n_roof, "cells_full_matrix_per_step": cells_step,
                    "gcups_computed_cells_kernel_only": ws * 32 / dp_s / 1e9 if dp_s else None,
                    "traceback": {"kernel": "fxg::walk2_kernel<4> (one lane per alignment, recomputes the tiles its path crosses)",
                                  "kernel_ms_per_step": roof_ctr["trace_kernel_ms"] / n_roof,
                                  "checkpoint_bytes_per_step": roof_ctr["trace_bytes"] / n_roof,
                                  "note": "latency-bound chain of dependent steps per alignment, not a bandwidth kernel"},
                    "single_stream_device_ms_per_step": roof_ctr["run_ms"] / n_roof,
                    "as_benchmarked": {
                        "note": "one batch on one stream with window sharing on (the production path): launches too small to fill 148 SMs "
                                "on their own, run concurrently with those of the other batches in flight",
                        "root_launch_frac": (prod_ctr["root_launch_word_steps"] * MYERS_INSTR_PER_WORD_STEP / (prod_ctr["root_launch_ms"] * 1e-3) / int32_peak)
                        if int32_peak and prod_ctr["root_launch_ms"] > 0 else None,
                        "root_launch_ms": prod_ctr["root_launch_ms"] / n_roof,
                        "root_launch_word_steps": prod_ctr["root_launch_word_steps"] / n_roof,
                        "all_launches_frac": (prod_ctr["dp_word_steps"] * MYERS_INSTR_PER_WORD_STEP / (prod_ctr["dp_kernel_ms"] * 1e-3) / int32_peak)
                        if int32_peak and prod_ctr["dp_kernel_ms"] > 0 else None,
                        "dp_word_steps_per_step": prod_ctr["dp_word_steps"] / n_roof,
                        "dp_kernel_ms_per_step": prod_ctr["dp_kernel_ms"] / n_roof,
                        "cells_computed_per_step": prod_ctr["dp_word_steps"] * 32 / n_roof,
                        "traceback_kernel_ms_per_step": prod_ctr["trace_kernel_ms"] / n_roof,
                        "checkpoint_bytes_per_step": prod_ctr["trace_bytes"] / n_roof,
                        "shared_score_passes_per_step": prod_ctr["shared_score_passes"] / n_roof,
                        "rescored_roots_per_step": prod_ctr["rescored_roots"] / n_roof,
                        "inferred_inner_per_step": prod_ctr["inferred_inner"] / n_roof,
                        "shared_tracebacks_per_step": prod_ctr["shared_tracebacks"] / n_roof,
                        "single_stream_device_ms_per_step": prod_ctr["run_ms"] / n_roof}}
        cpu_sec, cpu_stats, n_used = cpu_arm(refs, batch, cfg, args.cpu_sample_reads, threads)
        cpu_cells = cpu_stats["cells_inner"] + cpu_stats["cells_root"]
        line = {"metric": "pex_verification_gcups", "value": gcups, "unit": "GCUPS", "reads_per_s": reads_all / (ms_per_step * 1e-3),
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": config,
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"]},
                "e2e": {"value": e2e_gcups, "unit": "GCUPS", "reads_per_s": reads_all / (e2e_ms_per_step * 1e-3),
                        "ms_per_step": e2e_ms_per_step,
                        "h2d_bytes_per_step": e2e_ctr["h2d_bytes"] // n_e2e, "d2h_bytes_per_step": e2e_ctr["d2h_bytes"] // n_e2e},
                "gpu_launches": int(ctr["kernel_launches"]),
                "roofline": roofline,
                "cpu_baseline": {"value": cpu_cells / cpu_sec / 1e9, "unit": "GCUPS", "cores": threads, "kind": "port",
                                 "sample": f"first {n_used} reads of rank 0's batch, all anchors, both strands, CIGARs",
                                 "reads_per_s": n_used / cpu_sec},
                "step_latency_ms": [round(x, 3) for x in lat_ms], "step_latency_device_ms": [round(x, 3) for x in kernel_ms],
                "host_cores": len(os.sched_getaffinity(0)),
                "waves_per_step": ctr["waves"] / args.steps,
                "alignments_per_step": n_alignments,
                "stats_per_step": stats}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def _measured_peak(key: str):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


if __name__ == "__main__":
    sys.exit(main())
