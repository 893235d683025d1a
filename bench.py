#!/usr/bin/env python
"""Benchmark of the echelonization hot path (BASELINE.json: echelonize/rank time on the config-2 shape).

    python bench.py --gpus N --steps K --warmup W                 # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W # the reference's own CPU path (oracle/_ref)

One "step" = one spasm_echelonize of the workload (what tools/rank times, reference: tools/rank.c:100-103).
The JSON line carries:
  value / ms_per_step : CUDA-event time of a step with the matrix already resident in HBM (max over ranks)
  e2e                 : the same step through the reference-facing C ABI (spasm_echelonize on host buffers:
                        host->device copy of A, device->host copy of the echelon form U inside the timed region)
  roofline            : the dominant kernel of the step, timed live with CUDA events inside the library
  cpu_baseline        : oracle/_ref (the reference's C sources) on the host cores, rank 0, N=1
  schur               : "Schur rows/s" of SURVEY.md 8d: spasm_schur forced on the round-0 non-pivotal rows of config 1,
                        timed alone on the GPU and (N=1) on oracle/_ref
  scale_leg           : the workload whose work SHARDS (BASELINE config 4: echelonize + RREF + kernel basis of the
                        500k x 500k matrix, rows of U / free columns sharded, variable-length NCCL exchange), timed at
                        this N and, on rank 0 alone before the communicator exists, on one GPU: the 1 -> N speed-up of
                        the sharded path is then readable from a single line
  first_call_s        : the first spasm_echelonize of the process (CUDA context, pools, pinned staging)
`--workload all` prints one line per BASELINE config, config 2 first.
Nothing here reads /root/reference; oracle/ is used only as the CPU baseline / reference arm (in a subprocess: the
timed GPU process never imports it).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "echelonize/rank time (s)"
_libc = C.CDLL(None)


def reset_rand() -> None:
    """glibc rand() is part of the reference's result and never seeded there (seed 1)"""
    _libc.srand(1)


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "source": "fallback"}


def make_workload(name: str, scale: float):
    from spasm_b200 import synthetic
    if name == "config2":
        t = synthetic.config2(scale).transposed()      # tools/rank.c:84-88 transposes when n < m
        opts = {}
    elif name == "config1":
        t, opts = synthetic.config1(scale), {}
    elif name == "config3":
        t, opts = synthetic.config3(scale), {"sparsity_threshold": 0.01}
    elif name == "config4":
        t, opts = synthetic.config4(scale), {}
    elif name == "config5":
        t, opts = synthetic.config5(scale), {}
    else:
        raise SystemExit(f"unknown workload {name}")
    return t, opts


class ClockSampler:
    """SM clock / throttle reasons inside the timed region (B200_PROFILING.md recipe), sampled BETWEEN the timed steps:
    right after a step's last kernel, before the next one starts.  Every step is timed on its own (CUDA events inside
    the library, perf_counter around the C call for e2e), so a sample never sits inside a measured interval.  Sampling
    concurrently was tried three ways and perturbed the measurement each time: forking nvidia-smi from this process
    (GBs of mappings: the page-table copy stalled the timed thread ~20 ms), NVML from a thread, and NVML from a helper
    process (the queries take driver locks that CUDA calls wait for: +35 to +150 ms on the step they hit)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz, self.h = index, [], set(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()] if vis else []
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if index < len(ids) else index)
            self.bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                         pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            self.get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def sample(self):
        if os.environ.get("SPASM_B200_BENCH_NO_CLOCKS"):
            return
        try:
            if self.h is not None:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = self.get(self.h)
                self.reasons.update(n for n, b in zip(self.NAMES, self.bits) if mask & b)
                return
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            self.samples.append(float(out[0]))
            self.max_mhz = float(out[1])
            self.reasons.update(n for n, v in zip(self.NAMES, out[2:6]) if v.strip().lower() == "active")
        except Exception:
            pass

    def summary(self) -> dict:
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "when": "between timed steps"}


def run_reference(args) -> None:
    """The reference's own CPU implementation (oracle/_ref/libspasm_ref.so = /root/reference/src compiled in place,
    dense stage restated -- not FFLAS-FFPACK), all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import oracle
    from spasm_b200 import host
    t, opts = make_workload(args.workload, args.scale)
    R = oracle.ref()
    kind = "reference"
    if R is None:
        kind = "port"
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)                      # the reference prints \r progress bars on stderr
    try:
        times = []
        rank_found = None
        if R is not None:
            A = host.compress(R, t)
            for step in range(args.warmup + args.steps):
                reset_rand()
                t0 = time.perf_counter()
                f = host.echelonize(R, A, host.default_opts(R, **opts))
                dt = time.perf_counter() - t0
                rank_found = f.rank
                if step >= args.warmup:
                    times.append(dt)
        else:
            A = oracle.compress(t)
            cores = 1
            for step in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                e = oracle.echelonize(A, oracle.default_opts(**opts))
                dt = time.perf_counter() - t0
                rank_found = e.rank
                if step >= args.warmup:
                    times.append(dt)
    finally:
        os.dup2(saved, 2)
    v = sum(times) / len(times)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 mod p (balanced)", "data": "synthetic",
            "config": {"workload": f"{args.workload} (scale {args.scale}): {t.n}x{t.m}, {t.nz} entries, p={t.prime}", "rank": rank_found},
            "cpu_baseline": {"value": v, "unit": "s", "cores": cores, "kind": kind,
                             "sample": "the full workload, every step (dense stage: restated, not FFLAS-FFPACK)"},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def schur_leg(lib, host_mod, reps: int = 3):
    """SURVEY.md 8d "Schur rows/s": spasm_schur on the round-0 non-pivotal rows of config 1 (the default options never
    reach it: the Schur complement of every uniform-random config is dense).  Returns (rows, nnz, seconds per call)."""
    import numpy as np
    from spasm_b200 import abi
    t, _ = make_workload("config1", 1.0)
    A = host_mod.compress(lib, t)
    reset_rand()
    npiv, p, fact = host_mod.pivots_extract_structural(lib, A)
    rows = np.ascontiguousarray(p[npiv:], np.int32)
    p_out = np.zeros(max(len(rows), 1), np.int32)
    best, nnz = None, 0
    for _ in range(reps):
        t0 = time.perf_counter()
        S = host_mod.CsrHandle(lib, lib.spasm_schur(A.ptr, abi.as_int_p(rows), len(rows), C.byref(fact), 1.0, None, None, abi.as_int_p(p_out)))
        dt = time.perf_counter() - t0
        nnz = S.nnz
        del S
        best = dt if best is None else min(best, dt)
    return len(rows), int(nnz), best


def ingest_leg(lib) -> dict:
    """spasm_triplet_load on the SMS text of the headline workload, entry lines parsed by the host loop and by
    the GPU (csrc/gpu/ingest.cu); seconds include reading the file and the copy of the triplets to host memory."""
    import tempfile
    from spasm_b200 import synthetic
    t = synthetic.config2(1.0)
    text = t.to_sms()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    out = {"workload": f"SMS text of config2: {len(text)} bytes, {len(t.i)} entries"}
    with tempfile.NamedTemporaryFile(suffix=".sms") as tmp:
        tmp.write(text)
        tmp.flush()
        for mode in ("host", "gpu"):
            os.environ["SPASM_B200_INGEST"] = mode
            best = None
            for _ in range(3):
                f = libc.fopen(tmp.name.encode(), b"r")
                t0 = time.perf_counter()
                T = lib.spasm_triplet_load(f, t.prime, None)
                dt = time.perf_counter() - t0
                libc.fclose(f)
                assert T.contents.nz == len(t.i)
                lib.spasm_triplet_free(T)
                best = dt if best is None else min(best, dt)
            out[mode + "_s"] = best
            out[mode + "_MBps"] = len(text) / best / 1e6
        os.environ.pop("SPASM_B200_INGEST", None)
    return out


def run_schur_reference() -> None:
    """the same call through oracle/_ref on all host cores (subprocess of the GPU arm, rank 0, N=1)"""
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle
    from spasm_b200 import host
    R = oracle.ref()
    if R is None:
        print(json.dumps({"ref_seconds": None}))
        return
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)
    try:
        rows, nnz, sec = schur_leg(R, host, reps=1)
    finally:
        os.dup2(saved, 2)
    print(json.dumps({"ref_seconds": sec, "rows": rows, "nnz": nnz, "cores": os.cpu_count() or 1}))


def cpu_baseline(args, t, opts) -> dict:
    """Bounded CPU sample on the host cores: one full run of the workload through oracle/_ref."""
    cores = os.cpu_count() or 1
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--workload", args.workload, "--scale", str(args.scale)]
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), RANK="0", WORLD_SIZE="1")
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env).stdout.strip().splitlines()
        d = json.loads(out[-1])
        return d["cpu_baseline"]
    except Exception as exc:   # the baseline is reported, never required
        return {"value": None, "unit": "s", "cores": cores, "kind": "reference", "sample": f"failed: {exc}"}


def scale_leg(L, host, spasm_b200, barrier, steps: int, scale: float, reduce_max=None) -> dict:
    """BASELINE config 4: echelonize + RREF + kernel basis; with several ranks the rows of U (RREF) and the non-pivotal
    columns (kernel) are sharded and the result rows exchanged over NCCL (csrc/gpu/api.cu: exchange_pieces)."""
    t, opts = make_workload("config4", scale)
    A = host.compress(L, t)
    o = host.default_opts(L, **opts)
    out = {"ech": [], "rref": [], "kernel": [], "total": [], "nccl": 0, "rank": None, "rref_nnz": None, "kernel_dim": None}
    for step in range(steps + 1):                # one untimed pass first
        barrier()
        reset_rand()
        L.spasm_b200_reset_stats()
        t0 = time.perf_counter()
        f = host.echelonize(L, A, o)
        barrier()                                # a barrier after every call: each component is the time of its slowest rank
        t1 = time.perf_counter()
        Rm, _ = host.rref(L, f)
        barrier()
        t2 = time.perf_counter()
        Km = host.kernel(L, f)
        barrier()
        t3 = time.perf_counter()
        if step > 0:
            s = spasm_b200.Stats()
            L.spasm_b200_get_stats(C.byref(s))
            out["ech"].append(t1 - t0); out["rref"].append(t2 - t1); out["kernel"].append(t3 - t2); out["total"].append(t3 - t0)
            out["nccl"] = int(s.nccl_bytes)
        out["rank"], out["rref_nnz"], out["kernel_dim"] = f.rank, int(Rm.nnz), int(Km.n)
        del Rm, Km, f
    res = {k: sum(out[k]) / len(out[k]) for k in ("ech", "rref", "kernel", "total")}
    if reduce_max is not None:
        res = reduce_max(res)
    res.update(workload=f"config4 (scale {scale}): {t.n}x{t.m}, {t.nz} entries, p={t.prime}: echelonize + rref + kernel through the C ABI",
               rank=out["rank"], rref_nnz=out["rref_nnz"], kernel_dim=out["kernel_dim"], nccl_bytes_per_step=out["nccl"], steps=steps)
    return res


def bench_one(args, workload: str, L, host, spasm_b200, torch, dist, sampler, first_call) -> dict | None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    t, opts = make_workload(workload, args.scale)
    A = host.compress(L, t)
    o = host.default_opts(L, **opts)
    handle = L.spasm_b200_upload_csr(A.ptr)
    ms = C.c_double()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps
    for _ in range(max(args.warmup, 3)):
        reset_rand()
        L.spasm_b200_echelonize_resident(handle, C.byref(o), C.byref(ms))
    barrier()
    step_ms, agg = [], None
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        L.spasm_b200_flush_l2()              # cold L2 between timed iterations (the workload, ~10 MB, would fit in L2)
        reset_rand()
        L.spasm_b200_reset_stats()
        rk = L.spasm_b200_echelonize_resident(handle, C.byref(o), C.byref(ms))
        step_ms.append(ms.value)
        sampler.sample()                     # right after the step's last kernel; outside every measured interval
        s = spasm_b200.Stats()
        L.spasm_b200_get_stats(C.byref(s))
        if agg is None:
            agg = {k: 0.0 for k in ("ms_pivots", "ms_pivots_greedy", "ms_solve", "ms_dense", "ms_dense_gemm", "ms_k_greedy",
                                    "ms_k_panel_solve", "kernel_launches", "greedy_edges", "solve_traffic_model", "gemm_fieldops",
                                    "solve_rows", "solve_bytes", "gemm_int8_ops", "nccl_bytes")}
        for k in agg:
            agg[k] += float(getattr(s, k))
    barrier()
    wall = time.perf_counter() - wall0
    total_ms = sum(step_ms)

    # ---- end-to-end steps through the reference-facing C ABI (host buffers in, host echelon form out)
    e2e_times, h2d, d2h = [], 0, 0
    reset_rand()
    host.echelonize(L, A, o)                 # warm-up
    barrier()
    for _ in range(args.steps):
        L.spasm_b200_flush_l2()
        reset_rand()
        L.spasm_b200_reset_stats()
        t0 = time.perf_counter()
        f = host.echelonize(L, A, o)
        e2e_times.append(time.perf_counter() - t0)
        s = spasm_b200.Stats()
        L.spasm_b200_get_stats(C.byref(s))
        h2d, d2h = int(s.h2d_bytes), int(s.d2h_bytes)
        del f
    barrier()
    e2e_total = sum(e2e_times)
    L.spasm_b200_free_csr(handle)

    if dist is not None:
        v = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(v[0]), float(v[1])
    if rank != 0:
        return None
    K = args.steps
    ms_per_step = total_ms / K
    # with N > 1 the ranks cooperate on ONE echelonization (strong scaling): the time of the slowest rank
    value = ms_per_step / 1e3
    pk = peaks()
    per = {k: v / K for k, v in agg.items()}
    # per-kernel roofline table: algorithmic bytes / live CUDA-event time of that kernel in a step.
    # dram traffic: only from an `ncu --set full` capture of THIS workload (profiles/traffic.json, keyed by workload)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(workload, {})
    table = {
        "greedy_pivot_search": {"ms": per["ms_k_greedy"], "bytes": 4.0 * per["greedy_edges"],
                                "note": "4 B x pivot-row entries traversed by the searches (counted on the device)"},
        "panel_solve": {"ms": per["ms_k_panel_solve"], "bytes": per["solve_bytes"],
                        "note": "per solved row: 8 B x entries of the pivotal rows it reaches + 4 B x dense output columns (SURVEY 8d)"},
    }
    for k, v in table.items():
        v["GBps"] = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0.0
        v["frac"] = v["GBps"] / pk["hbm_gbs"]
        v["traffic"] = traffic.get(k)      # dram bytes of one step's launches of that kernel family, or null
    dominant = max(table, key=lambda k: table[k]["ms"])
    d = table[dominant]
    kernels = {k: v["ms"] for k, v in table.items()}
    kernels["dense_echelon"] = per["ms_dense"]
    int8_peak = peaks_int8()
    line = {
        "metric": METRIC, "value": value, "unit": "s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 mod p (balanced)", "data": "synthetic",
        "config": {"workload": f"{workload} (scale {args.scale}): {t.n}x{t.m}, {t.nz} entries, p={t.prime}",
                   "rank": int(rk), "l2": "flushed between timed steps (512 MB write)",
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} GPUs cooperate on one matrix: rows of every dense block sharded + ncclAllGather; pivot search and dense echelon replicated"
                   " (the echelonization of this shape is depth-bound and does not shard: see scale_leg for the path that does)",
                   "value_excludes": "assemble(): dense rows -> CSR and the download of the echelon form (inside e2e)"},
        "e2e": {"value": e2e_total / K, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "first_call_s": first_call,
        "gpu_launches": int(round(per["kernel_launches"])) * K,
        "clocks": sampler.summary(),
        "roofline": {"kernel": dominant, "bound": "hbm", "achieved": d["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": d["GBps"] / pk["hbm_gbs"], "traffic": d["traffic"], "peak_source": pk["source"], "bytes_per_launch": d["bytes"],
                     "ms_per_launch": kernels[dominant], "note": d["note"]},
        "kernels": {k: {"ms_per_step": round(v["ms"], 3), "algorithmic_GBps": round(v["GBps"], 2), "frac_of_hbm_peak": round(v["frac"], 5),
                        "traffic": v["traffic"]} for k, v in table.items()},
        "phases_ms": {k: round(per[k], 3) for k in ("ms_pivots", "ms_pivots_greedy", "ms_k_greedy", "ms_solve", "ms_k_panel_solve",
                                                    "ms_dense", "ms_dense_gemm")},
        "dense_modp_tops": per["gemm_fieldops"] / (per["ms_dense"] / 1e3) / 1e12 if per["ms_dense"] > 0 else None,
        "dense_int8_tensor_ops_per_step": per["gemm_int8_ops"],
        "dense_int8_frac_of_measured_peak": (per["gemm_int8_ops"] / (per["ms_dense_gemm"] / 1e3) / 1e12 / int8_peak["tops"]
                                             if per["ms_dense_gemm"] > 0 and int8_peak["tops"] else None),
        "int8_peak": int8_peak,
        "nccl_bytes_per_step": per["nccl_bytes"],
        "wall_s_timed_region": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        sub_args = argparse.Namespace(**vars(args))
        sub_args.workload = workload
        line["cpu_baseline"] = cpu_baseline(sub_args, t, opts)
    return line


def peaks_int8() -> dict:
    """measured dense int8 tensor peak (tools/int8_peak.py -> profiles/int8_peak.json); sustained figure: the products
    run inside a long step"""
    path = os.path.join(ROOT, "profiles", "int8_peak.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tops": float(d["int8_tops_sustained"]), "burst_tops": float(d["int8_tops_burst"]), "source": "measured (cuBLASLt int8, profiles/int8_peak.json)"}
    return {"tops": None, "burst_tops": None, "source": "unmeasured"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "schur-reference"])
    ap.add_argument("--workload", default="config2", help="config1..config5, or `all` (one JSON line per config, config2 first)")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale-leg", action="store_true")
    ap.add_argument("--no-schur-leg", action="store_true")
    ap.add_argument("--scale-leg-steps", type=int, default=2)
    ap.add_argument("--scale-leg-scale", type=float, default=1.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.impl == "schur-reference":
        run_schur_reference()
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local_rank)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("SPASM_B200_DEVICE", str(local_rank))

    import torch
    import spasm_b200
    from spasm_b200 import host

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = spasm_b200.lib()          # raises if the CUDA library was not built: no fallback
    L.spasm_b200_set_verbose(0)
    workloads = ["config2", "config1", "config3", "config4", "config5"] if args.workload == "all" else [args.workload]

    # ---- the first call of the process (one-shot tools/rank pays this once)
    t0w, o0w = make_workload(workloads[0], args.scale)
    A0 = host.compress(L, t0w)
    reset_rand()
    t0 = time.perf_counter()
    f0 = host.echelonize(L, A0, host.default_opts(L, **o0w))
    first_call = time.perf_counter() - t0
    del f0, A0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- scale leg, single-GPU figure: rank 0 alone, before the library's communicator exists
    leg1 = None
    want_leg = not args.no_scale_leg and args.workload in ("config2", "all")
    if want_leg and world > 1:
        if rank == 0:
            leg1 = scale_leg(L, host, spasm_b200, lambda: torch.cuda.synchronize(), 1, args.scale_leg_scale)
        dist.barrier()

    if dist is not None:
        # the library's own NCCL communicator: solve batches, rows of U (rref) and free columns (kernel) are sharded
        from spasm_b200 import sharding
        sharding.init_comm(L, dist, device=torch.device("cuda", local_rank))

    lines = []
    for w in workloads:
        lines.append(bench_one(args, w, L, host, spasm_b200, torch, dist, sampler, first_call))

    # ---- the leg whose work shards (BASELINE config 4)
    legN = None
    if want_leg:
        def reduce_max(res):
            if dist is None:
                return res
            keys = sorted(res)
            v = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            return {k: float(x) for k, x in zip(keys, v)}
        legAll = None
        if world > 1:
            # first with every rank receiving the full result (one group of ncclBroadcast), then gathered on rank 0 only
            # (ncclSend / ncclRecv; the other ranks skip a download they do not use): the second is the leg's value
            legAll = scale_leg(L, host, spasm_b200, barrier, args.scale_leg_steps, args.scale_leg_scale, reduce_max)
            L.spasm_b200_comm_result_root(0)
        legN = scale_leg(L, host, spasm_b200, barrier, args.scale_leg_steps, args.scale_leg_scale, reduce_max)
        if world > 1:
            L.spasm_b200_comm_result_root(-1)

    if rank == 0:
        if legN is not None:
            leg = {"value": legN["total"], "unit": "s", "n_gpus": world, "workload": legN["workload"],
                   "echelonize_s": legN["ech"], "rref_s": legN["rref"], "kernel_s": legN["kernel"], "steps": legN["steps"],
                   "rank": legN["rank"], "rref_nnz": legN["rref_nnz"], "kernel_dim": legN["kernel_dim"],
                   "nccl_bytes_per_step": legN["nccl_bytes_per_step"],
                   "collective": ("ncclAllGather (piece sizes) + one group of ncclSend / ncclRecv, one per result piece (variable length), to rank 0: "
                                  "the result (a host matrix) is materialised on rank 0" if world > 1 else "none (one rank)"),
                   "sharding": "rows of U (rref) and non-pivotal columns (kernel) in contiguous slices, U replicated; echelonize: rows of the solve batches",
                   "timing": "host wall clock around the three C-ABI calls (they return host matrices), barrier on both sides, max over ranks"}
            if legAll is not None:
                leg["every_rank_gets_the_result"] = {"value": legAll["total"], "rref_s": legAll["rref"], "kernel_s": legAll["kernel"],
                                                     "nccl_bytes_per_step": legAll["nccl_bytes_per_step"],
                                                     "collective": "one group of ncclBroadcast, one per result piece, from its owner; every rank downloads the full host matrix"}
            single = leg1 if leg1 is not None else (legN if world == 1 else None)
            if single is not None:
                leg["single_gpu_value"] = single["total"]
                leg["single_gpu_rref_s"] = single["rref"]
                leg["speedup_vs_single_gpu"] = single["total"] / legN["total"]
                leg["rref_speedup_vs_single_gpu"] = single["rref"] / legN["rref"]
            lines[0]["scale_leg"] = leg
        if not args.no_schur_leg and world == 1 and args.workload in ("config2", "all"):
            try:
                rows, nnz, sec = schur_leg(L, host)
                sch = {"workload": "config1 (20000x20000): spasm_schur forced on the round-0 non-pivotal rows (SURVEY 8d)", "rows": rows, "nnz_out": nnz,
                       "seconds": sec, "rows_per_s": rows / sec, "timing": "best of 3 C-ABI calls, host arrays in and out"}
                if world == 1 and not args.no_cpu_baseline:
                    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "schur-reference"]
                    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
                    ref = json.loads(out[-1])
                    if ref.get("ref_seconds"):
                        sch["reference_rows_per_s"] = ref["rows"] / ref["ref_seconds"]
                        sch["reference_cores"] = ref["cores"]
                lines[0]["schur"] = sch
                lines[0]["schur_rows_per_s"] = sch["rows_per_s"]
            except Exception as exc:     # a leg, never the headline
                lines[0]["schur"] = {"failed": str(exc)}
            try:
                lines[0]["ingest"] = ingest_leg(L)
            except Exception as exc:
                lines[0]["ingest"] = {"failed": str(exc)}
        for line in lines:
            print(json.dumps(line), flush=True)
    if dist is not None:
        L.spasm_b200_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
