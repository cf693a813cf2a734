#!/usr/bin/env python
"""bench.py — agent-updates/sec of the per-frame agent update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--agents A]

One "step" = one frame (bin -> forces+seek -> walls -> integrate -> triangle walk) over every agent.
N=1 workload: BASELINE config 3 — 1M agents, all MOVING in 64 command groups that follow precomputed
funnel paths on the 1024x1024-cell constrained-Delaunay map with 200 buildings.
Prints ONE JSON line (rank 0).  See the task contract in DESIGN.md §Measurement.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 112   # persistent per-agent record read once + written once


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~5 ms from a thread (the timed
    region is tens of milliseconds, an `nvidia-smi -lms` child would not deliver its first row in time); nvidia-smi once
    as the fallback when NVML cannot be loaded."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.rows = []          # (sm_mhz, reasons bitmask)
        self.index = index
        self.max_mhz = None
        self.h = None
        self.nv = None
        self.stop = threading.Event()
        self.t = None
        self.period_s = float(os.environ.get("RTS_CLOCK_POLL_MS", "5")) * 1e-3
        try:
            self._open()
        except Exception:
            self.nv = None

    def _open(self):
        import pynvml as nv
        nv.nvmlInit()
        try:
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            self.h = nv.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        self.nv = nv

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            why = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            why = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.rows.append((sm, why))

    def _poll(self):
        while not self.stop.is_set():
            try:
                self._sample()
            except Exception:
                return
            time.sleep(self.period_s)

    def __enter__(self):
        # NVML is opened in the constructor: nvmlInit takes tens of milliseconds and, with 8 ranks on one host, a different
        # time in every rank; inside the timed region that start-up skew shows up as time spent waiting for the neighbour
        try:
            if self.nv is None:
                self._open()
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None
        return self

    def __exit__(self, *a):
        if self.nv is not None:
            try:
                self._sample()          # the work is still in flight or has just ended: at least one row under load
            except Exception:
                pass
            self.stop.set()
            if self.t:
                self.t.join(timeout=1)
        else:
            self._smi_once()

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            c = [x.strip() for x in out.split(",")]
            why = 0
            for k, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if c[2 + k].lower().startswith("active"):
                    why |= self.REASONS[nm]
            self.rows.append((float(c[0]), why))
            self.max_mhz = float(c[1])
        except Exception:
            pass

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[1]
        reasons = sorted(nm for nm, bit in self.REASONS.items() if mask & bit)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def build_world(n_agents, seed=1236):
    from rtsengine_b200 import scenario
    fx = scenario.load_map("map_c3.npz")
    agents, paths = scenario.config3_agents(fx, n=n_agents, seed=seed)
    return fx, scenario.tables_of(fx), agents, paths


def pinned(shape, dtype):
    import torch
    t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t, t.numpy()


def cpu_port_baseline(tables, agents, paths, budget_s=20.0):
    """oracle port (same source as the parity oracle, reference Release flags) on all host cores, bounded sample"""
    from oracle import port
    n = len(agents["r"])
    P = port.default_params(1024, 1024, fast=True)
    ow = port.OracleWorld(P, tables, paths, fast=True)
    names = {f[0] for f in __import__("rtsengine_b200").abi.AGENT_FIELDS if f[0] != "gid"}
    a = __import__("rtsengine_b200").abi.alloc_agent_arrays(n, names)
    for k, v in agents.items():
        a[k][:] = v
    ow.step(a)   # warm-up frame
    t0 = time.perf_counter()
    steps = 0
    while steps < 2 or (time.perf_counter() - t0 < budget_s and steps < 50):
        ow.step(a)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "agent-updates/s", "cores": int(port.load(True).rts_oracle_threads()),
            "kind": "port", "sample": f"{n} agents x {steps} frames of the same workload (oracle/rts_oracle.c, -Ofast, OpenMP)"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from rtsengine_b200 import abi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from rtsengine_b200 import slabs
        return slabs.bench_multi_gpu(args, measured_peak, ClockSampler, ALGO_BYTES_PER_UPDATE)

    n = args.agents
    fx, tables, agents, paths = build_world(n)
    params = engine.default_params(1024, 1024)
    g = engine.GpuAgentSystem(params, device=local)
    g.upload_map(tables)
    g.upload_paths(paths)
    g.set_agents(agents)

    K, Wm = args.steps, max(args.warmup, 3)
    g.update(Wm, sync=True)

    # ---- device-resident throughput: K frames (CUDA graph of 8 frames replayed), CUDA events on the launching stream ----
    g.reset_timings()
    with ClockSampler(local) as clocks:
        g.sync()
        g.update(K)
        g.sync()
    t = g.timings()
    total_ms = t.last_step_ms
    launches = int(t.kernel_launches)
    value = n * K / (total_ms * 1e-3)

    # ---- per-kernel durations over THE SAME frames: same initial state, same warm-up, then K frames launched one by
    #      one with a CUDA-event pair around every kernel (crowd density, hence kernel time, depends on the frame index) ----
    g.set_agents(agents)
    g.update(Wm, sync=True)
    g.set_profiling(True)
    g.reset_timings()
    g.update(K, sync=True)
    kt = {k: g.kernel_time_ms(k) for k in ("step", "dense", "scan", "reorder", "walk")}
    g.set_profiling(False)
    step_ms = kt["step"][0]
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_UPDATE * n / (step_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if int(tj.get("agents", 0)) == n:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "k_step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * n,
                "kernel_ms": {k: v[0] for k, v in kt.items()},
                "whole_frame_frac": ALGO_BYTES_PER_UPDATE * n / (total_ms / K * 1e-3) / 1e9 / peak}

    # ---- end to end through the System2-shaped API with HOST buffers (pinned): per frame
    #      updateSharedData (H2D blackboard) -> update -> communicate (D2H blackboard) ----
    #      The blackboard (SharedData, ECS.h:136-145) lives in pinned host memory; the same buffers are uploaded and
    #      written back, as the reference's systems read and write entity2shared_data in place.
    g.set_agents(agents)
    g.update(Wm, sync=True)
    board = {"r": pinned((n, 2), np.float32), "vel": pinned((n, 2), np.float32), "angle": pinned((n,), np.float32),
             "angle_vel": pinned((n,), np.float32), "state": pinned((n,), np.uint8), "r_next": pinned((n, 2), np.float32)}
    cur = g.communicate(("r", "angle", "angle_vel", "state"))
    for k in cur:
        board[k][1][:] = cur[k]
    up_ptrs = {k: board[k][1].ctypes.data for k in ("r", "angle", "angle_vel", "state")}
    down_ptrs = {k: v[1].ctypes.data for k, v in board.items()}
    h2d = sum(board[k][1].nbytes for k in up_ptrs)
    d2h = sum(v[1].nbytes for v in board.values())

    def e2e_frame():
        g.update_shared_data_ptrs(up_ptrs)     # System2::updateSharedData   (H2D)
        g.update(1)                            # System2::update
        g.communicate_ptrs(down_ptrs)          # System2::communicate        (D2H, synchronises)

    for _ in range(3):
        e2e_frame()
    Ke = min(K, 100)
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_frame()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * Ke / e2e_s, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
           "api": "GpuAgentSystem.update_shared_data -> update -> communicate (rts_update_agents/rts_step/rts_download), pinned host buffers"}
    g.close()

    cpu = None
    if not args.no_cpu:
        try:
            cpu = cpu_port_baseline(tables, agents, paths, budget_s=args.cpu_budget)
        except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "agent-updates/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}

    line = {
        "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": 1, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "config3: 1M agents, 64 groups on precomputed funnel paths, 1024x1024-cell CDT map with 200 buildings",
                   "agents": n, "map_cells": [1024, 1024], "buildings": 200, "groups": 64,
                   "l2_policy": f"inputs larger than L2: {n} agents x ~190 B touched per frame across ping-pong buffers "
                                f"({n * 190 / 1e6:.0f} MB vs 126 MB L2); no explicit flush"},
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


def run_reference(args):
    """The reference's own CPU implementation (verbatim TUs, Release flags, all host threads) on a bounded sample
    of the same workload: 5 000 agents (the reference's cap) at config-3 density on a 512² window with buildings."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is meant to use every host thread, and the
    # OpenMP runtime reads the variable when the library is loaded
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import refh
    K, Wm = args.steps, max(args.warmup, 3)
    n_world = int(os.environ.get("WORLD_SIZE", "1"))
    if refh.available(fast=True):
        L = refh.load(True)
        L.refh_set_omp_threads.argtypes = [__import__("ctypes").c_int]
        L.refh_set_omp_threads(os.cpu_count() or 1)   # (in case an OpenMP runtime was initialised before the variable was set)
        rng = np.random.default_rng(1236)
        w = refh.RefWorld(512, 512, fast=True)
        placed = 0
        occ = np.zeros((512, 512), bool)
        while placed < 50:   # same building density as config 3 (200 per 1024² cells)
            cx, cy = (int(v) for v in rng.integers(12, 500, 2))
            nn, mm = (8, 8) if rng.random() < 0.5 else (4, 6)
            x0, y0 = cx - nn // 2, cy - mm // 2
            if occ[x0 - 2:x0 + nn + 2, y0 - 2:y0 + mm + 2].any():
                continue
            if w.add_building(cx, cy, nn, mm, 1) != 0:
                continue
            occ[x0 - 1:x0 + nn + 1, y0 - 1:y0 + mm + 1] = True
            placed += 1
        w.finalize_map()
        n = 5000
        side = float(np.sqrt(n / (1_000_000 / 5120.0 ** 2)))   # window with config-3 density
        pos = []
        while len(pos) < n:
            p = rng.uniform(1000.0, 1000.0 + side, 2)
            if not occ[int(p[0] // 5), int(p[1] // 5)]:
                pos.append(p)
        for p in pos:
            w.add_agent(float(p[0]), float(p[1]))
        w.issue_paths(np.arange(n, dtype=np.int32), 2300.0, 2200.0)
        w.step(Wm)
        t0 = time.perf_counter()
        w.step(K)
        dt = time.perf_counter() - t0
        threads = int(refh.load(True).refh_omp_threads())
        kind, sample = "reference", (f"{n} agents (reference cap N_MAX_NAVIGABLE_BOIDS) at config-3 density in a {side:.0f}² window, 50 buildings, "
                                     f"verbatim reference TUs -Ofast, {K} frames")
        value = n * K / dt
    else:
        from oracle import port
        port.load(True).rts_oracle_set_threads(os.cpu_count() or 1)
        fx, tables, agents, paths = build_world(200_000)
        cpu = cpu_port_baseline(tables, agents, paths, budget_s=20.0)
        value, threads, kind, sample = cpu["value"], cpu["cores"], "port", cpu["sample"]
        dt = None
    line = {"impl": "reference", "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": n_world,
            "steps": K, "warmup": Wm, "ms_per_step": (1e3 * dt / K) if dt else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config3: 1M agents, 64 groups on precomputed funnel paths, 1024x1024-cell CDT map with 200 buildings"},
            "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--agents", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
