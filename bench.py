#!/usr/bin/env python
"""bench.py -- dyn_core cell-updates/s of the FV3 acoustic-dynamics hot path on B200.

Contract (see task statement / BASELINE.json):
  python bench.py --gpus N --steps K --warmup W            -> our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  -> the reference's CPU path
                                                              (C++ oracle port, all host cores)
One "step" of our arm = one dyn_core call (n_split acoustic substeps) over the full cubed sphere
(6 faces, halo exchanges included).  metric = nx*ny*nz*n_split*ntiles / t.
N=1: all 6 faces of C384L79 on one GPU (device-local halo gathers); N>1: faces are spread over
min(N,6) ranks (one process per GPU, NCCL P2P for off-rank faces) -> "strong" scaling: total
work is fixed at the full C384L79 cube.
The reference arm runs the SAME configuration (C384L79 full cube, same flags, same initial state) on the host cores;
one of its steps is a bounded sample of the workload: a dyn_core call of --ref-substeps (default 1) acoustic substeps
(every substep costs the same), counted as that many cell updates.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dyn_core cell-updates/sec (nx*ny*nz*n_split/s) at C384L79; d_sw HBM GB/s"

from gfdl_atmos_cubed_sphere_b200.parallel import tiles_of_rank  # noqa: E402


def dsw_dram_traffic(n, npz, flagset):
    """dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of ONE d_sw call on one face, from the committed
    ncu metrics list of this round (profiles/r2_dsw_traffic.json, written by profiles/dsw_traffic.py --json from the csv of
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ... profiles/prof_dsw.py`; the per-kernel
    breakdown stays in that file).  None for a workload without a committed capture."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_dsw_traffic.json")))
        ent = rec.get(f"C{n}L{npz}{flagset}")
        if not ent:
            return None, None
        return float(ent["bytes"]), {k: ent[k] for k in ("read", "write", "launches", "ncu_serialised_ms", "csv") if k in ent}
    except Exception:
        return None, None


def dsw_algorithmic_bytes(n, npz, use_cond=False, d_con=False):
    """compulsory d_sw traffic per face, all levels (bytes)."""
    R = 3 * (n + 6) ** 2 + 2 * (n + 6) * (n + 7) + 2 * (n + 7) * (n + 6) + (n + 7) ** 2 + 2 * (n + 1) * n + 2 * (n + 1) * (n + 6)
    W = 3 * n * n + 2 * n * (n + 1) + 4 * (n + 1) * (n + 6) + 2 * n * (n + 1) + 2 * (n + 1) * (n + 6)
    if use_cond:
        R += (n + 6) ** 2; W += n * n
    if d_con:
        R += n * n; W += n * n
    return 8 * npz * (R + W)


class ClockSampler:
    def __init__(self, device=0):
        self.samples, self.reasons, self.stop = [], set(), False
        self.device = device
        self.max_mhz = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.th.start()

    def finish(self):
        self.stop = True
        self.th.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_case(n, npz, flagset="A"):
    from gfdl_atmos_cubed_sphere_b200 import Case
    return Case(n, npz, flagset, state="baroclinic")


def workload_config(args):
    """Identical in both arms: the configuration the metric is quoted on."""
    return {"workload": f"C{args.res}L{args.npz} nonhydrostatic full cube (6 faces), n_split={args.n_split}, flag-set {args.flagset}, "
                        f"JW baroclinic wave, set_eta L79 levels (var_hi, ptop 1 Pa), dt_atmos={args.dt_atmos}s",
            "l2": "working set per stage (>= 6 fields x 96 MB per face) exceeds the 126 MB L2; no explicit flush",
            "n_split": args.n_split, "faces": 6}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gfdl_atmos_cubed_sphere_b200 import abi, CudaCube

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, npz, n_split = args.res, args.npz, args.n_split
    bdt = args.dt_atmos
    my_tiles = tiles_of_rank(rank, world)
    case = build_case(n, npz, args.flagset)
    lib = abi.load_library()
    cube = CudaCube.for_rank(case, rank, world, device=local)
    if world > 1:
        CudaCube.attach_nccl(cube, rank, world)   # the library's own NCCL communicator over the active ranks
    if args.transport_fp32 and cube is not None:
        cube.set_transport_fp32(True)             # BASELINE config 5 (not the headline): fp32 PPM sweeps on d_sw's interior tiles

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if cube is not None:
            cube.dyn_core(bdt, n_split, graph=args.graph)

    # pinned host copies of the prognostic state for the e2e (host-buffer) leg
    e2e_fields = ["U", "V", "W", "DELZ", "PT", "DELP", "PHIS"]
    pinned = {}
    if cube is not None:
        for t in my_tiles:
            for f in e2e_fields:
                a = cube.eng[t].get(f)
                p = torch.empty(a.shape, dtype=torch.float64).pin_memory()
                p.numpy()[...] = a
                pinned[(t, f)] = p

    def step_e2e():
        h2d = d2h = 0
        if cube is None:
            return 0, 0
        for t in my_tiles:
            e = cube.eng[t]
            for f in e2e_fields:
                p = pinned[(t, f)]
                e.check(e._fn("put_field")(e.ctx, abi.FIELD_ID[f], C.c_void_p(p.data_ptr())), "put_field")
                h2d += p.numel() * 8
        cube.dyn_core(bdt, n_split, graph=args.graph)
        for t in my_tiles:
            e = cube.eng[t]
            for f in ("U", "V", "W", "DELZ", "PT", "DELP"):
                p = pinned[(t, f)]
                e.check(e._fn("get_field")(e.ctx, abi.FIELD_ID[f], C.c_void_p(p.data_ptr())), "get_field")
                d2h += p.numel() * 8
        return h2d, d2h

    def reset_state():
        if cube is None:
            return
        for t in my_tiles:
            case.load_state(cube.eng[t], t)

    for _ in range(max(args.warmup, 3)):
        step()
    # the state is re-initialised so the timed steps integrate the same physical regime
    reset_state()
    launches0 = sum(lib[0].fv3_launch_count(cube.eng[t].ctx) for t in my_tiles) if cube else 0
    if cube is not None:
        for t in my_tiles:
            lib[0].fv3_stage_timers(cube.eng[t].ctx, 1)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    # CUDA events on the library's own launch streams (torch events only see torch's stream)
    if cube is not None:
        lib[0].fv3_timer_start(cube.ctxs, len(my_tiles))
    for _ in range(args.steps):
        step()
    t_wall = 0.0
    if cube is not None:
        ms = C.c_double(0)
        lib[0].fv3_timer_stop(cube.ctxs, len(my_tiles), C.byref(ms))
        t_wall = ms.value / 1e3
    barrier()
    clocks = sampler.finish() if sampler else None
    stage_ms = {}
    launches = 0
    if cube is not None:
        launches = sum(lib[0].fv3_launch_count(cube.eng[t].ctx) for t in my_tiles) - launches0
        for nm in ("C_SW", "UPDATE_DZ_C", "Riem_Solver_C", "PG_C", "D_SW", "UPDATE_DZ", "Riem_Solver3", "PG_D"):
            ms_tot, calls_tot = 0.0, 0
            for t in my_tiles:
                ms, calls = C.c_double(0), C.c_longlong(0)
                lib[0].fv3_stage_time_ms(cube.eng[t].ctx, nm.encode(), C.byref(ms), C.byref(calls))
                ms_tot += ms.value; calls_tot += calls.value
            stage_ms[nm] = (ms_tot, calls_tot)
        for t in my_tiles:
            lib[0].fv3_stage_timers(cube.eng[t].ctx, 0)
    tt = torch.tensor([t_wall], dtype=torch.float64, device="cuda")
    ln = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)   # kernels launched by ALL ranks inside the timed region
    t_max = float(tt.item())
    launches = int(ln.item())
    # ---- roofline leg: the dominant stage (d_sw, batched over k) on ONE face with nothing else in
    # flight, CUDA events on its launch stream.  (The per-stage timers above overlap the 6 face streams,
    # so they give shares of the step, not kernel durations.)
    dsw_solo_ms = None
    if rank == 0 and cube is not None:
        reset_state()
        torch.cuda.synchronize()
        e0 = cube.eng[my_tiles[0]]
        one = (C.c_void_p * 1)(e0.ctx)
        dts = bdt / n_split
        e0.call("c_sw", 0.5 * dts)
        for _ in range(3):
            e0.call("d_sw", dts)
        e0.sync()
        reps = 5
        lib[0].fv3_timer_start(one, 1)
        for _ in range(reps):
            e0.call("d_sw", dts)
        ms = C.c_double(0)
        lib[0].fv3_timer_stop(one, 1, C.byref(ms))
        dsw_solo_ms = ms.value / reps
    # ---- e2e leg (host buffers through the C ABI, H2D + D2H inside the timed region)
    reset_state()
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        a, b = step_e2e()
        h2d, d2h = a, b
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    hb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    if rank == 0:
        cells = n * n * npz * n_split * 6
        value = cells * args.steps / t_max
        e2e_value = cells * e2e_steps / float(te.item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6650 GB/s)"
        alg = dsw_algorithmic_bytes(n, npz, bool(case.flags.get("use_cond")), case.flags.get("d_con", 0) > 1e-5)
        achieved = (alg / 1e9) / (dsw_solo_ms / 1e3) if dsw_solo_ms else None
        traffic, traffic_src = dsw_dram_traffic(n, npz, args.flagset)
        cfg = workload_config(args)
        cfg["faces_per_rank"] = len(tiles_of_rank(0, world))
        cfg["launch"] = "CUDA graph (FV3_DYN_GRAPH)" if args.graph and world == 1 else "direct kernel launches"
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (fp32 PPM sweeps on d_sw interior tiles: --transport-fp32)" if args.transport_fp32 else "f64", "data": "synthetic",
            "config": cfg,
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": int(hb[0].item()),
                    "d2h_bytes_per_step": int(hb[1].item())},
            "roofline": {"bound": "hbm", "kernel": "d_sw (all kernels of one batched-over-k d_sw call on one face)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": which, "algorithmic_bytes_per_launch": alg,
                         "ms_per_launch": dsw_solo_ms,
                         # secondary ceiling: fp64 CUDA-core throughput.  FLOP model: 1.1 kFLOP per cell and d_sw call
                         # (4-5 fv_tp_2d + 2 momentum PPM sweeps + damping); peak = DFMA rate measured with
                         # profiles/micro/fp64_rate.cu on this pool's B200 (60.4 lanes/clk/SM x 148 SMs x 1.965 GHz x 2)
                         "fp64": {"flop_per_cell": 1100, "achieved_tflops": (1100.0 * n * n * npz / (dsw_solo_ms / 1e3) / 1e12) if dsw_solo_ms else None,
                                  "peak_tflops": 35.1, "peak_source": "measured DFMA rate, profiles/r1_fp64_rate_b200.txt"},
                         "launch": "one fv3_d_sw call = every kernel of d_sw for all npz levels of one face"},
            "stage_ms_per_call": {k: (v[0] / v[1] if v[1] else None) for k, v in stage_ms.items()},
            "clocks": clocks,
        }
        if args.cpu_baseline:
            # bounded sample of the SAME workload: 2 acoustic substeps of the C384L79 cube on all host cores (~10-25 s)
            del pinned
            line["cpu_baseline"] = cpu_reference(args, case, steps=2, warmup=1, budget_s=40.0)[0]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_reference(args, case=None, steps=5, warmup=1, budget_s=200.0):
    """The oracle port (timing build: -O3, OpenMP over k / j like the reference's own threading, automatic arrays from a
    per-thread stack, 6-tile halo exchange inside the library) on ALL host cores, on the configuration of the metric.
    One step = one dyn_core call of --ref-substeps acoustic substeps (bounded sample: every substep costs the same).
    Returns (cpu_baseline dict, executed steps, seconds per step)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H   # test infrastructure: the one place bench.py executes oracle/
    case = case or build_case(args.res, args.npz, args.flagset)
    lib = H.load_oracle(fast=True)[0]
    cores = int(lib.fv3o_use_all_cores())    # torchrun exports OMP_NUM_THREADS=1 to its workers: take every core anyway
    oc = H.OracleCube(case, fast=True)
    ns = max(1, args.ref_substeps)
    bdt = args.dt_atmos * ns / args.n_split   # same acoustic time step as the GPU arm
    t0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        oc.dyn_core(bdt, ns)                  # first touch of every array, thread pool start
    t_warm = (time.perf_counter() - t0) / max(1, warmup)
    steps_exec = int(max(1, min(steps, budget_s // max(t_warm, 1e-3))))
    timers = {}
    t0 = time.perf_counter()
    for _ in range(steps_exec):
        oc.dyn_core(bdt, ns, timers)
    wall = time.perf_counter() - t0
    oc.close()
    cells = args.res * args.res * args.npz * ns * 6
    stage_s = sum(timers.values())
    cb = {"value": cells * steps_exec / wall, "unit": "cell-updates/s", "cores": cores, "kind": "port",
          "sample": f"full cube C{args.res}L{args.npz} (the configuration of the metric), {steps_exec} dyn_core call(s) of {ns} acoustic "
                    f"substep(s) each after {max(1, warmup)} warm-up call(s); C++ oracle port, timing build (-O3 -march=x86-64-v3 "
                    f"-fopenmp, halo exchange inside the library), {cores} OpenMP threads",
          "wall_s": wall, "steps": steps_exec, "substeps_per_step": ns,
          "halo_and_driver_share": round(1.0 - stage_s / wall, 4),
          "stage_seconds_per_step": {k: round(v / steps_exec, 4) for k, v in timers.items()}}
    return cb, steps_exec, wall / steps_exec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, steps_exec, s_per_step = cpu_reference(args, steps=args.steps, warmup=max(1, min(args.warmup, 2)), budget_s=200.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "cell-updates/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps_exec, "steps_requested": args.steps,
            "warmup": max(1, min(args.warmup, 2)), "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference Fortran cannot be built here (no Fortran compiler, FMS not vendored): this arm times the C++ "
                    "restatement (oracle port) on all host cores; one step = a bounded sample (ref-substeps acoustic substeps) of "
                    "the same C384L79 workload, steps = the number of steps actually executed"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=384)
    ap.add_argument("--npz", type=int, default=79)
    ap.add_argument("--n-split", dest="n_split", type=int, default=8)
    ap.add_argument("--dt-atmos", dest="dt_atmos", type=float, default=225.0)
    ap.add_argument("--flagset", default="A")
    ap.add_argument("--ref-substeps", dest="ref_substeps", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--graph", action="store_true",
                    help="fv3_dyn_core(FV3_DYN_GRAPH): the step as one CUDA graph (single-process runs; bit-identical to the direct launches)")
    ap.add_argument("--transport-fp32", dest="transport_fp32", action="store_true",
                    help="mixed precision (BASELINE config 5): fv3_set_transport_fp32; the default headline run is all fp64")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
