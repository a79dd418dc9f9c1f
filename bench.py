#!/usr/bin/env python
"""bench.py -- headline benchmark of the HS/HP hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, one process per GPU under torchrun)
  python bench.py --impl reference [...]                       the reference's own CPU algorithm (oracle port)

A step = one pass of the hot path over one batch: 65,536 random flop situations PER GPU (BASELINE.json configs[2];
weak scaling, the situations of the batch are independent), HS + full HP enumeration in Treys mode through
torch.ops.holdem.hp_out -> holdem_hp_batch; N > 1: each rank computes its block and the kernels store every result row into
every rank's symmetric-memory buffer over NVLink (fused compute + all-gather; NCCL all-gather where that is unavailable).
Every rank then recomputes the whole batch and compares.  Prints ONE JSON line on rank 0 (the only line on stdout).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dask-holdem-simulator_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

SITUATIONS_PER_GPU = 65536
SEED = 20251019
EVALS_PER_SITUATION = 1_072_353          # 1,070,190 + 1,081 + 1,082
HOT_KERNEL = "hp_flop_treys_v4_kernel"
METRIC = "flop_hs_hp_situations_per_sec"
UNIT = "situations/s"


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- CPU arm: the reference's algorithm on the host cores ----------------------------------------------------------------
def _cpu_chunk(args):
    sit_bytes, n, mode = args
    import numpy as np
    from oracle import oracle as O
    sit = np.frombuffer(sit_bytes, dtype=np.uint8).reshape(n, 8)
    return O.hp_batch(sit, mode).tobytes()


def cpu_port_throughput(n_situations, cores, repeats=1):
    """Oracle port (oracle/ehs_oracle.c: the reference's HS/HP loops with Treys' evaluator, C, one process per core)
    on the first n_situations of the bench workload.  Returns (situations/s, seconds, result rows)."""
    import multiprocessing as mp
    import numpy as np
    import holdem_b200 as hb
    from oracle import oracle as O
    O.build()
    sit = hb.random_situations(SITUATIONS_PER_GPU, SEED)[:n_situations]
    chunks = [sit[i::cores] for i in range(cores)]
    chunks = [c for c in chunks if len(c)]
    best = None
    rows = None
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        pool.map(_cpu_chunk, [(np.ascontiguousarray(c[:1]).tobytes(), 1, "treys") for c in chunks])  # warm the workers
        for _ in range(repeats):
            t0 = time.perf_counter()
            parts = pool.map(_cpu_chunk, [(np.ascontiguousarray(c).tobytes(), len(c), "treys") for c in chunks])
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best = dt
            rows = np.zeros((n_situations, 16), dtype=np.uint32)
            for i, p in enumerate(parts):
                rows[i::cores] = np.frombuffer(p, dtype=np.uint32).reshape(-1, 16)
    return n_situations / best, best, rows


def python_path_evals_per_sec(n_hands=20000):
    """Pure-Python Treys restatement (oracle/py_treys.py), one core: 7-card evaluate() calls per second."""
    import random
    from oracle.py_treys import PyEvaluator, card_from_path
    ev = PyEvaluator()
    rng = random.Random(0)
    hands = [[card_from_path(c) for c in rng.sample(range(52), 7)] for _ in range(n_hands)]
    t0 = time.perf_counter()
    for h in hands:
        ev.evaluate(h[:2], h[2:])
    return n_hands / (time.perf_counter() - t0)


CONFIG = {
    "workload": "configs[2]: 65,536 random flop situations per GPU, HS + full HP enumeration, mode=treys "
                "(numpy default_rng(20251019) deals)",
    "situations_per_gpu": SITUATIONS_PER_GPU, "mode": "treys",
    "l2": "GPU arm: 256 MiB buffer written between timed iterations (L2 flush), inputs 0.5 MiB/GPU; CPU arm: not applicable",
}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_step = max(32, 4 * cores)   # bounded sample of the 65,536-situation workload per step (~1 s of CPU work)
    for _ in range(args.warmup):
        cpu_port_throughput(min(per_step, cores), cores)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt, _ = cpu_port_throughput(per_step, cores)
        t_total += dt
    value = per_step * args.steps / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {per_step} situations of the workload per step x {args.steps} steps, "
                                   "oracle/ehs_oracle.c (reference loop structure: one opponent evaluation per (holding, "
                                   "runout) pair, Treys min-over-21 evaluator, our rank cached per runout), one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hand_evals_per_sec": value * EVALS_PER_SITUATION,
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:
        return {}


# ---- our arm -----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the evaluator / micro-benchmark / other-config side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import holdem_b200 as hb
    from holdem_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG is left as the driver set it (its log is how the driver counts ranks) but the log is routed to stderr, so
        # that stdout holds exactly one line: the JSON (NCCL still prints a line at process exit, after anything we could print).
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    hb._lib.init(local_rank)
    golden = load_json("tests", "golden", "gpu_hashes.json")

    def all_ranks(ok: bool) -> bool:
        """AND over ranks of a local verdict."""
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: 65,536 situations per GPU; rank r owns rows [r*65536, (r+1)*65536) of the global batch ---------------
    n_local = SITUATIONS_PER_GPU
    n_global = n_local * world
    sit_all = hb.random_situations(n_global, SEED)           # same on every rank (seeded)
    sit_np = np.ascontiguousarray(sit_all[rank * n_local:(rank + 1) * n_local])
    d_sit = torch.from_numpy(sit_np).to(dev)
    d_sit_all = torch.from_numpy(sit_all).to(dev) if world > 1 else d_sit
    full = torch.zeros((n_global, hb.HP_ROW), dtype=torch.int32, device=dev)
    mine = full[rank * n_local:(rank + 1) * n_local]
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(fused=None):
        """One pass of the hot path over the global batch; every rank ends up with all rows.  N > 1: each rank computes its
        contiguous block and the kernels store the rows into every rank's symmetric-memory buffer over NVLink (compute and
        all-gather fused in one launch); falls back to an in-place NCCL all-gather where symmetric memory is unavailable."""
        if world == 1:
            hb.hand_potential_counts(d_sit, "treys", out=mine)        # torch.ops.holdem.hp_out -> holdem_hp_batch
            return full
        return hb.hand_potential_counts_sharded(d_sit_all, "treys", fused=fused)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    fused_used = world > 1 and sharded.fused_path_error(dev) is None

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = hb.kernel_launches()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush_buf.fill_(1)          # L2 flush between timed iterations (not timed)
        e0.record()
        result = step()
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches_timed = hb.kernel_launches() - launches0
    dev_ms = max_over_ranks(sum(e0.elapsed_time(e1) for e0, e1 in evs))
    ms_per_step = dev_ms / args.steps
    value = n_global / (ms_per_step * 1e-3)
    if result is not full:
        full.copy_(result)

    # ---- correctness of what was timed: EVERY rank recomputes EVERY row of the global batch on its own GPU and compares it
    #      with the buffer the timed path delivered (its own rows and the rows the other GPUs stored / gathered) ----------------
    verify = {}
    local_all = hb.hand_potential_counts(d_sit_all, "treys") if world > 1 else None
    if world > 1:
        verify["timed_path_rows_equal_single_gpu_recompute_on_every_rank"] = all_ranks(torch.equal(full, local_all))
    allgather_value = None
    if fused_used:      # the same step with the NCCL all-gather instead of the kernels' own peer stores, for comparison
        for _ in range(3):
            got = step(fused=False)
        barrier()
        verify["nccl_allgather_rows_equal_single_gpu_recompute_on_every_rank"] = all_ranks(torch.equal(got, local_all))
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            step(fused=False)
        a1.record()
        barrier()
        allgather_value = n_global / (max_over_ranks(a0.elapsed_time(a1)) / args.steps * 1e-3)
    # content hash (position-dependent, 64 bit) of the gathered buffer: equal on every rank, and its first 65,536 rows equal to
    # the committed single-GPU hash (tests/golden/gpu_hashes.json; those rows were compared with the C oracle one by one)
    hash_global = hb.rows_hash(full)
    hash_first = hb.rows_hash(full[:SITUATIONS_PER_GPU])
    if world > 1:
        t = torch.tensor([hash_global >> 32, hash_global & 0xFFFFFFFF], dtype=torch.int64, device=dev)
        t0_ = t.clone()
        dist.broadcast(t0_, src=0)
        verify["hash_equal_on_every_rank"] = all_ranks(torch.equal(t, t0_))
    want_hash = golden.get("configs2_first_65536_rows_treys")
    verify["hash_first_65536_rows_equals_committed"] = None if want_hash is None else (f"{hash_first:016x}" == want_hash)

    # ---- kernel-only time of the dominant kernel (same launches, no collective) for the roofline ----------------------------
    k_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in k_evs:
        flush_buf.fill_(1)
        e0.record()
        hb.hand_potential_counts(d_sit, "treys", out=mine)
        e1.record()
    torch.cuda.synchronize()
    kernel_ms = sum(e0.elapsed_time(e1) for e0, e1 in k_evs) / args.steps
    clocks = sampler.stop()

    # ---- e2e: host buffers in, host buffers out, through the C ABI's host entry point (copies inside the timed region) --------
    # Page-locked caller buffers (the contract's "pinned host memory") are copied from / to directly; pageable ones go through
    # the library's own pinned staging (timed as well, reported beside it).  N > 1: every rank serves its own block of the
    # global batch from / to its own host buffers (the result of the job ends up in host memory, one block per process).
    def time_host_path(h_in, h_out, steps):
        hb.hand_potential_counts_host(h_in, "treys", device=local_rank, out=h_out)   # warm (allocates buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            hb.hand_potential_counts_host(h_in, "treys", device=local_rank, out=h_out)
        torch.cuda.synchronize()
        return n_global * steps / max_over_ranks(time.perf_counter() - t0)

    e2e_steps = max(3, args.steps // 2)
    pin_in = torch.empty((n_local, 8), dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = sit_np
    pin_out = torch.zeros((n_local, hb.HP_ROW), dtype=torch.int32, pin_memory=True)
    host_out = pin_out.numpy().view(np.uint32)
    e2e_value = time_host_path(pin_in.numpy(), host_out, e2e_steps)
    pageable_out = np.empty((n_local, hb.HP_ROW), dtype=np.uint32)
    e2e_pageable = time_host_path(sit_np, pageable_out, e2e_steps)

    rows = full.cpu().numpy().astype(np.int64)
    my_rows = rows[rank * n_local:(rank + 1) * n_local]
    verify["row_invariants"] = all_ranks(bool((rows[:, 12:15].sum(axis=1) == 1081).all()
                                              and (rows[:, 3:12].reshape(-1, 3, 3).sum(axis=2) == 990 * rows[:, 12:15]).all()))
    verify["host_entry_rows_equal_device_rows"] = all_ranks(np.array_equal(host_out.astype(np.int64), my_rows)
                                                            and np.array_equal(pageable_out, host_out))

    # ---- configs[4]: 1,000,000 random flop situations, strong scaling over the N GPUs (side key, every N) -----------------------
    c5 = None
    if not args.no_extras:
        n5 = 1_000_000
        sit5 = torch.from_numpy(hb.random_situations(n5, seed=20251021)).to(dev)
        for _ in range(2):
            r5 = hb.hand_potential_counts_sharded(sit5, "treys")
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps5 = 5
        a0.record()
        for _ in range(reps5):
            r5 = hb.hand_potential_counts_sharded(sit5, "treys")
        a1.record()
        barrier()
        ms5 = max_over_ranks(a0.elapsed_time(a1)) / reps5
        r5 = r5.contiguous()
        h5 = hb.rows_hash(r5)
        want5 = golden.get("configs4_1m_flop_rows_treys")
        c5 = {"situations": n5, "n_gpus": world, "ms": ms5, "situations_per_sec": n5 / (ms5 * 1e-3), "scaling": "strong",
              "content_hash": f"{h5:016x}", "hash_equals_committed_single_gpu_hash": None if want5 is None else f"{h5:016x}" == want5,
              "what": "configs[4]: 1M random flop situations (seed 20251021) HS+HP sharded over the GPUs, every rank receives all "
                      "rows; the hash is position-dependent over all 16 M result words and must be the same for N = 1/2/4/8"}
        if world > 1:   # every rank recomputes all 1M rows itself and compares
            c5["rows_equal_single_gpu_recompute_on_every_rank"] = all_ranks(
                torch.equal(r5, hb.hand_potential_counts(sit5, "treys")))
            verify["config5_rows_equal_single_gpu_recompute_on_every_rank"] = c5["rows_equal_single_gpu_recompute_on_every_rank"]
        if want5 is not None:
            verify["config5_hash_equals_committed"] = all_ranks(c5["hash_equals_committed_single_gpu_hash"])
        del sit5, r5

    ok = all(v for v in verify.values() if v is not None)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- rank 0 only: roofline denominators, side measurements, CPU baseline ---------------------------------------------------
    peaks = load_json("MEASURED_PEAKS.json")
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = load_json("profiles", "traffic.json")

    def gather_rate(table_bytes, pattern):
        return hb.smem_gather_microbench(table_bytes=table_bytes, iters=4096, pattern=pattern, blocks_per_sm=1 if table_bytes > 65536 else 2,
                                         threads=1024, device=local_rank)[0]
    lookup_peak_random = gather_rate(131072, 0)          # production footprint (the evaluator's 7-card val table is 128 KB)
    lookup_peak_cf = gather_rate(131072, 1)
    lookup_peak_random_4k = gather_rate(4096, 0)
    counters = load_json("profiles", "hot_kernel_counters.json").get(HOT_KERNEL, {})
    evals_per_launch = n_local * EVALS_PER_SITUATION
    achieved_evals = evals_per_launch / (kernel_ms * 1e-3)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6
    issue_measured = {name: hb.int_issue_microbench(mix=mix, device=local_rank)[0]
                      for mix, name in ((0, "compare_plus_predicated_mad"), (1, "fma_pipe_only"), (2, "alu_pipe_only"))}
    smem_peak = 148 * sm_mhz * 1e6
    wi = counters.get("warp_instructions_per_situation")
    wf = counters.get("shared_wavefronts_per_situation")
    issue_rate = None if wi is None else wi * n_local / (kernel_ms * 1e-3)
    roofline = {
        "kernel": HOT_KERNEL, "bound": "issue_slots",
        "achieved": None if issue_rate is None else issue_rate / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
        "frac": None if issue_rate is None else issue_rate / issue_peak,
        "traffic": (lambda t: t.get("dram_bytes_read", 0) + t.get("dram_bytes_write", 0) if t else None)(
            traffic.get(HOT_KERNEL)),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture "
                        "(profiles/traffic.json); algorithmic HBM bytes per launch = 65,536 x 72 B = 4.7 MB -- HBM is not the bound",
        "note": "The kernel is bound by instruction issue (and, close behind, shared-memory wavefronts), not by HBM or tensor cores "
                "(integer lookup work).  achieved = warp instructions per situation (smsp__inst_executed.sum / 65,536 of the "
                "committed ncu capture, profiles/hot_kernel_counters.json) x 65,536 / this run's CUDA-event kernel time; "
                "peak = 148 SMs x 4 schedulers x the SM clock sampled during this run.  The ALGEBRAIC ratio (1,072,353 "
                "evaluations per situation against the measured random-gather rate) is under algorithmic_evaluations: the "
                "kernel obtains those counts from a rank-multiset sum + flush corrections without visiting every pair "
                "(DESIGN.md section 4), so that ratio exceeds 1 and is not a utilisation",
        "kernel_ms": kernel_ms,
        "issue_slots": None if wi is None else {
            "warp_instructions_per_situation": wi,
            "achieved_warp_instr_per_sec": issue_rate,
            "peak_warp_instr_per_sec": issue_peak,
            "frac": issue_rate / issue_peak,
            "measured_issue_rates_warp_instr_per_sec": issue_measured,
            "frac_of_measured_mixed_issue_rate": issue_rate / issue_measured["compare_plus_predicated_mad"]},
        "shared_memory": None if wf is None else {
            "wavefronts_per_situation": wf,
            "achieved_wavefronts_per_sec": wf * n_local / (kernel_ms * 1e-3),
            "peak_wavefronts_per_sec": smem_peak,
            "frac": wf * n_local / (kernel_ms * 1e-3) / smem_peak,
            "note": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum / 65,536 from the same capture; peak = one wavefront per "
                    "SM per clock"},
        "algorithmic_evaluations": {
            "per_situation": EVALS_PER_SITUATION, "per_sec": achieved_evals,
            "random_gather_peak_per_sec": {"table_128KB": lookup_peak_random, "table_4KB": lookup_peak_random_4k,
                                           "table_128KB_conflict_free": lookup_peak_cf},
            "ratio_to_random_gather_peak_128KB": achieved_evals / lookup_peak_random,
            "hash_lookups_per_situation_actually_issued": 0,
            "note": "uniformly random uint16 shared-memory gathers from every lane of a full grid, measured in this run on a "
                    "table of the production footprint (128 KB) and on a 4 KB table"},
        "ncu": {k: counters.get(k) for k in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "lsu_pipe_pct",
                                              "shared_wavefront_pct", "capture")} if counters else None,
        "hbm_algorithmic_GBs": n_local * 72 / (kernel_ms * 1e-3) / 1e9, "hbm_peak_GBs": hbm_peak,
        "hbm_peak_source": hbm_peak_src,
    }

    # ---- the other implementations of the same counts, for scale (kernel-only, CUDA events) ---------------------------------
    def _time(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    alt = {}
    extras = {}
    if not args.no_extras:
        scratch = torch.empty_like(mine)
        ms3 = _time(lambda: hb.hand_potential_counts(d_sit, "treys", out=scratch, flop_variant=3), 3)
        same3 = bool(torch.equal(scratch, mine))
        hash3 = hb.rows_hash(scratch)
        n_plain = 2048
        d_plain = d_sit[:n_plain].contiguous()
        plain_out = torch.empty((n_plain, hb.HP_ROW), dtype=torch.int32, device=dev)
        msp = _time(lambda: hb.hand_potential_counts(d_plain, "treys", out=plain_out, direct=True), 3)
        samep = bool(torch.equal(plain_out, mine[:n_plain]))
        ok = ok and same3 and samep
        alt = {
            "four_set_walk_v3": {"situations_per_sec": n_local / (ms3 * 1e-3), "equal_rows": same3,
                                 "content_hash": f"{hash3:016x}",
                                 "what": "deduplicating 4-set kernel (178,365 evaluations + 1.07 M compare steps per situation)"},
            "plain_enumeration": {"situations_per_sec": n_plain / (msp * 1e-3), "situations": n_plain, "equal_rows": samep,
                                  "evals_per_sec": n_plain * EVALS_PER_SITUATION / (msp * 1e-3),
                                  "frac_of_lookup_peak": n_plain * EVALS_PER_SITUATION / (msp * 1e-3) / lookup_peak_random,
                                  "what": "hp_direct_kernel: one real evaluation (shared-memory table gathers) per (holding, "
                                          "runout) pair -- the literal shape of the reference loop"},
        }
        # mode=reference (the reference's code as written: 9-category ranker, 1176 runouts) on the same 65,536 situations
        ref_out = torch.empty((n_local, hb.HP_ROW), dtype=torch.int32, device=dev)
        msr = _time(lambda: hb.hand_potential_counts(d_sit, "reference", out=ref_out), 5)
        big_hs = torch.from_numpy(hb.random_situations(1 << 20, seed=SEED + 1)).to(dev)
        msh = _time(lambda: hb.hand_strength_counts(big_hs, "reference"), 5)
        msh_t = _time(lambda: hb.hand_strength_counts(big_hs, "treys"), 5)
        n_old = 4096
        old_out = torch.empty((n_old, hb.HP_ROW), dtype=torch.int32, device=dev)
        mso = _time(lambda: hb.hand_potential_counts(d_sit[:n_old].contiguous(), "reference", out=old_out, ref_variant=4), 2)
        same_old = bool(torch.equal(old_out, ref_out[:n_old]))
        plain_ref = hb.hand_potential_counts(d_plain, "reference", direct=True)
        same_plain = bool(torch.equal(plain_ref, ref_out[:n_plain]))
        ok = ok and same_old and same_plain
        rc = load_json("profiles", "hot_kernel_counters.json").get("hp_ref5_kernel<flop>", {})
        wi_r = rc.get("warp_instructions_per_situation")
        ref_hash = f"{hb.rows_hash(ref_out):016x}"
        extras["mode_reference"] = {
            "hp_flop_situations_per_sec": n_local / (msr * 1e-3), "hp_situations": n_local, "ms": msr,
            "hs_situations_per_sec": big_hs.shape[0] / (msh * 1e-3), "hs_situations": int(big_hs.shape[0]),
            "hs_treys_mode_situations_per_sec": big_hs.shape[0] / (msh_t * 1e-3),
            "content_hash": ref_hash,
            "hash_equals_committed": None if "configs2_first_65536_rows_reference" not in golden
                                     else ref_hash == golden["configs2_first_65536_rows_reference"],
            "equal_to_earlier_4set_kernel": same_old, "earlier_4set_kernel_situations_per_sec": n_old / (mso * 1e-3),
            "equal_to_plain_enumeration": same_plain,
            "roofline": None if wi_r is None else {
                "kernel": "hp_ref5_kernel", "bound": "issue_slots (ALU pipe 86 % busy in the committed capture)",
                "achieved": wi_r * n_local / (msr * 1e-3) / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instructions/s",
                "frac": wi_r * n_local / (msr * 1e-3) / issue_peak, "warp_instructions_per_situation": wi_r,
                "ncu": {k: rc.get(k) for k in ("issue_active_pct", "alu_pipe_pct", "fma_pipe_pct", "shared_wavefront_pct", "capture")}},
            "what": "HOLDEM_MODE_REFERENCE: HandCalculations.py:45-82 / :178-226 exactly as written (golden-pinned to the "
                    "unmodified reference in tests/): rank-multiset kernel hp_ref5_kernel, rank-pair HS kernel hs_ref5_kernel"}
        ok = ok and extras["mode_reference"]["hash_equals_committed"] is not False

    def _eval_ncu_counters():
        c = load_json("profiles", "hot_kernel_counters.json").get("eval_batch_fast_kernel<7>")
        if not c:
            return None
        return {k: c.get(k) for k in ("warp_instructions_per_32_hands", "shared_wavefronts_per_32_hands",
                                      "lsu_data_pipe_wavefront_pct", "shared_wavefront_pct", "issue_active_pct",
                                      "alu_pipe_pct", "fmaheavy_pipe_pct", "capture")}

    if not args.no_extras:
        # batched 7-card evaluator fed from HBM: 10 algorithmic bytes per evaluation
        n_hands = 1 << 25
        hands = torch.empty((n_hands, 8), dtype=torch.uint8, device=dev)
        hands.fill_(255)
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        chunk = 1 << 21
        for s in range(0, n_hands, chunk):
            hands[s:s + chunk, :7] = torch.rand((chunk, 52), device=dev, generator=g).topk(7, dim=1).indices.to(torch.uint8)
        out16 = torch.empty(n_hands, dtype=torch.int16, device=dev)
        for _ in range(3):
            hb.evaluate(hands, 7, out=out16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            hb.evaluate(hands, 7, out=out16)
        e1.record()
        torch.cuda.synchronize()
        ev_ms = e0.elapsed_time(e1) / reps
        ev_rate = n_hands / (ev_ms * 1e-3)
        sample = hands[:200000].cpu().numpy()
        from oracle import oracle as O
        ev_ok = bool(np.array_equal(out16[:200000].cpu().numpy().astype(np.uint16), O.treys_evaluate_batch(sample, 7)))
        ok = ok and ev_ok
        extras["eval7_stream"] = {
            "evals_per_sec_per_gpu": ev_rate, "hands": n_hands, "ms": ev_ms, "verified_vs_oracle": ev_ok,
            "roofline": {"kernel": "eval_batch_fast_kernel<7>", "bound": "hbm",
                         "achieved": 10 * ev_rate / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": 10 * ev_rate / 1e9 / hbm_peak, "peak_source": hbm_peak_src,
                         "traffic": (lambda t: t.get("dram_bytes_read", 0) + t.get("dram_bytes_write", 0) if t else None)(
                             traffic.get("eval_batch_fast_kernel<7>")),
                         "note": "inputs 256 MiB + outputs 64 MiB per launch (> 126 MB L2); traffic from the committed ncu "
                                 "capture (profiles/traffic.json)",
                         "ncu": _eval_ncu_counters()}}
        del hands, out16
        # exhaustive sweep, hands generated in-kernel (BASELINE.json configs[1])
        hb.exhaustive7(local_rank)
        e0.record()
        hist, sums = hb.exhaustive7(local_rank)
        e1.record()
        torch.cuda.synchronize()
        ex_ms = e0.elapsed_time(e1)
        hist_hash = hb.rows_hash(hist.view(torch.int32).reshape(-1, 2))
        extras["eval7_exhaustive"] = {"hands": 133784560, "ms": ex_ms, "evals_per_sec_per_gpu": 133784560 / (ex_ms * 1e-3),
                                      "sum_rank_ok": sums.cpu().tolist() == [547965983972, 2665153084455708],
                                      "distinct_ranks": int((hist > 0).sum()),
                                      "rank_histogram_hash": f"{hist_hash:016x}",
                                      "rank_histogram_hash_equals_committed":
                                          None if "exhaustive7_rank_histogram" not in golden
                                          else f"{hist_hash:016x}" == golden["exhaustive7_rank_histogram"]}
        ok = ok and extras["eval7_exhaustive"]["sum_rank_ok"] and extras["eval7_exhaustive"]["rank_histogram_hash_equals_committed"] is not False

    # ---- configs[3]: 22-seat table, every seat on flop / turn / river, 100,000 hands (rank 0, N = 1) -------------------------
    if not args.no_extras and world == 1:
        from holdem_b200 import table
        from oracle import oracle as O
        n_hands_t = 100_000
        hole, board = table.deal_tables(n_hands_t, seed=20251020, seats=22)
        d_hole, d_board = torch.from_numpy(hole).to(dev), torch.from_numpy(board).to(dev)
        c4 = {"hands": n_hands_t, "seats": 22, "situations_per_street": n_hands_t * 22, "streets": {}}
        total_ms = 0.0
        rng = np.random.default_rng(4)
        c4_ok = True
        for nb, name, n_or in ((3, "flop_hs_hp", 768), (4, "turn_hs_hp", 4096), (5, "river_hs", 8192)):
            fn = lambda: hb.table_hand_potential_counts(d_hole, d_board, nb, "treys")
            ms = _time(fn, 2)
            total_ms += ms
            out = fn().reshape(-1, hb.HP_ROW)
            pick = np.sort(rng.choice(out.shape[0], n_or, replace=False))
            sit_pick = table.situations_for_street(hole, board, nb)[pick]
            same = bool(np.array_equal(out[torch.from_numpy(pick).to(dev)].cpu().numpy().astype(np.uint32),
                                       O.hp_batch_parallel(sit_pick, "treys", cache4=True)))
            h = hb.rows_hash(out)
            key = f"configs3_22seat_100k_{name}"
            same_hash = None if key not in golden else f"{h:016x}" == golden[key]
            c4_ok = c4_ok and same and same_hash is not False
            c4["streets"][name] = {"ms": ms, "situations_per_sec": out.shape[0] / (ms * 1e-3), "content_hash": f"{h:016x}",
                                   "hash_equals_committed": same_hash, "oracle_sample_rows": n_or, "oracle_sample_equal": same}
            del out
        c4["total_ms"] = total_ms
        c4["hands_per_sec"] = n_hands_t / (total_ms * 1e-3)
        c4["what"] = ("configs[3]: one shuffle per hand (seed 20251020), seat i holds cards 2i, 2i+1, board = cards 44..48; "
                      "holdem_table_hp_batch on flop, turn and river boards (river: hand strength only)")
        extras["config4_22_seat_table"] = c4
        ok = ok and c4_ok
        del d_hole, d_board

    cpu = None
    if world == 1:
        from oracle import oracle as O
        cores = host_cores()
        v0, dt0, _ = cpu_port_throughput(2 * cores, cores)            # probe, then size the sample for ~12 s of CPU work
        n_cpu = int(min(8192, max(64, 12.0 * v0)))
        v, dt, cpu_rows = cpu_port_throughput(n_cpu, cores)
        cpu_equal = bool(np.array_equal(cpu_rows.astype(np.int64), rows[:n_cpu]))   # GPU == oracle on the CPU sample
        # the oracle with its 4-set cache (same counts, ~6x fewer evaluations) checks a larger slice of what was timed
        n_chk = int(min(SITUATIONS_PER_GPU, max(256, 12.0 * v0 * 5)))
        t0 = time.perf_counter()
        chk_rows = O.hp_batch_parallel(sit_np[:n_chk], "treys", cache4=True)
        dt_chk = time.perf_counter() - t0
        chk_equal = bool(np.array_equal(chk_rows.astype(np.int64), rows[:n_chk]))
        ok = ok and cpu_equal and chk_equal
        py_rate = python_path_evals_per_sec()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n_cpu} situations of the same workload, {dt:.1f} s wall, oracle/ehs_oracle.c "
                         "(reference loop structure: one opponent evaluation per (holding, runout) pair, Treys min-over-21 "
                         "evaluator in C, our rank cached per runout), one process per core",
               "gpu_rows_equal_oracle_rows": cpu_equal,
               "hand_evals_per_sec": v * EVALS_PER_SITUATION,
               "oracle_4set_cache": {"rows_checked": n_chk, "seconds": dt_chk, "situations_per_sec": n_chk / dt_chk,
                                     "gpu_rows_equal_oracle_rows": chk_equal,
                                     "what": "same counts with the opponent's rank cached per 4-set of unseen cards "
                                             "(178,365 evaluations per situation); tests/test_gpu_hshp.py::"
                                             "test_hp_treys_all_65536_vs_oracle compares ALL rows of the workload this way"},
               "python_treys_path": {"evals_per_sec_per_core": py_rate,
                                     "situations_per_sec_per_core_est": py_rate / 2_141_462,
                                     "what": "pure-Python restatement of treys.Evaluator.evaluate (treys itself is "
                                             "not installable offline), 20,000 random 7-card hands, 1 core; a flop "
                                             "situation shaped like the reference loop makes 2,141,462 calls"}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": CONFIG,
        "global_situations": n_global,
        "parallelism": f"situation-sharded x{world}"
                       + ("" if world == 1 else (", result rows stored by the kernels into every rank's symmetric-memory buffer "
                                                 "over NVLink (fused compute + all-gather, one symmetric-memory barrier per step)"
                                                 if fused_used else ", in-place NCCL all-gather of the [N,16] u32 result")),
        "fused_path_error": sharded.fused_path_error(dev) if world > 1 else None,
        "nccl_allgather_path_value": allgather_value,
        "hand_evals_per_sec_per_gpu": value * EVALS_PER_SITUATION / world,
        "wall_s_timed_region": t_wall,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(sit_np.nbytes) * world,
                "d2h_bytes_per_step": int(host_out.nbytes) * world, "steps": e2e_steps,
                "api": "holdem_hp_batch_host (C ABI, page-locked caller buffers: one H2D copy of the rows, one launch, the kernels "
                       "store the result rows straight into the caller's mapped host buffer); N > 1: one such call per rank on its "
                       "own block, max over ranks",
                "pageable_caller_buffers": {"value": e2e_pageable, "unit": UNIT,
                                            "what": "same call with pageable numpy arrays: 16,384-row chunks through the library's "
                                                    "own pinned staging on two streams (two extra host memcpy per chunk)"}},
        "gpu_launches": int(launches_timed) * world,
        "gpu_launches_note": f"counted by the library (holdem_kernel_launches): {int(launches_timed)} launches on rank 0 inside the "
                             f"timed region of {args.steps} steps x {world} ranks; per step: hp_flop_treys_v4_kernel + "
                             "hp_turn_treys_v4_kernel + hs_treys_v4_kernel<river rows> (the last two exit at once when the batch "
                             "holds no turn / river rows)",
        "content_hash": {"global_rows": f"{hash_global:016x}", "first_65536_rows": f"{hash_first:016x}",
                         "committed_first_65536_rows": want_hash},
        "config5_1m_strong": c5,
        "other_implementations": alt,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "verify": verify,
        "verified": ok,
    }
    line.update(extras)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        time.sleep(1.0)             # let the other ranks' (and NCCL's) last log lines drain before the one JSON line
    sys.stdout.flush()
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
