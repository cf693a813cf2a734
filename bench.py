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


PRECISION_NAMES = {0: "exact", 1: "fast"}


def time_mode(g, agents, n, K, Wm, repeats, precision_name):
    """K frames from the SAME state (initial population + Wm warm-up frames), `repeats` times; median device time.
    Then the per-kernel durations over the same frames (profiling mode: kernels launched one by one, a CUDA-event pair
    around each — crowd density, hence kernel time, depends on the frame index)."""
    ms, launches = [], 0
    for _ in range(repeats):
        g.set_agents(agents)
        g.update(Wm, sync=True)
        g.reset_timings()
        g.update(K)
        g.sync()
        t = g.timings()
        ms.append(t.last_step_ms)
        launches = int(t.kernel_launches)
    g.set_agents(agents)
    g.update(Wm, sync=True)
    g.set_profiling(True)
    g.reset_timings()
    g.update(K, sync=True)
    kt = {k: g.kernel_time_ms(k)[0] for k in ("step", "dense", "scan", "reorder", "walk")}
    g.set_profiling(False)
    med = float(np.median(ms))
    return {"precision": precision_name, "value": n * K / (med * 1e-3), "ms_per_step": med / K, "repeat_ms_per_step": [m / K for m in ms],
            "launches": launches, "kernel_ms": kt}


def roofline_of(m, n, peak, peak_src, traffic):
    step_ms = m["kernel_ms"]["step"]
    achieved = ALGO_BYTES_PER_UPDATE * n / (step_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_step", "precision": m["precision"], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UPDATE * n, "kernel_ms": m["kernel_ms"],
            "whole_frame_frac": ALGO_BYTES_PER_UPDATE * n / (m["ms_per_step"] * 1e-3) / 1e9 / peak}


def k_step_traffic(n, precision_name):
    """DRAM bytes per k_step launch from the committed ncu capture of the same build (profiles/k_step_traffic.json)"""
    tp = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    try:
        tj = json.load(open(tp)).get(precision_name)
        if tj and int(tj.get("agents", 0)) == n:
            return tj["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def e2e_leg(g, n, agents, K, Wm):
    """The same metric through the System2-shaped calls with HOST buffers (pinned), per frame:
         updateSharedData : the entities the host touched since the last frame (here: a move command for 1 000 agents per
                            frame)                                           -> rts_update_agents_indexed   (H2D)
         update           : rts_step(1)
         communicate      : what the two systems write into SharedData — vel, angle_vel, state, target (21 B per agent,
                            PhysicsSystem.cpp:118-127, SeekSystem.cpp:322-342) — as a DIFFERENCE, the mirror image of the
                            upload: vel of every agent (it changes for everyone, 8 B) and {index, angle_vel, target, state} of
                            the entities for which one of the three changed (a few per cent per frame)
                                                                             -> rts_download_delta          (D2H, synchronises)
                            Before and after the timed frames the blackboard kept this way is compared bit for bit with a full
                            rts_download of the four fields.
       The loop's own integration r += vel*dt (ECS.cpp:98-113) reproduces the device's position bit for bit, so positions
       never travel (tests/test_gpu_adapter.py runs exactly this protocol inside the reference's frame loop)."""
    from rtsengine_b200 import abi
    g.set_agents(agents)
    g.update(Wm, sync=True)
    board = {"vel": pinned((n, 2), np.float32), "angle_vel": pinned((n,), np.float32), "state": pinned((n,), np.uint8),
             "r_next": pinned((n, 2), np.float32)}
    m = min(1000, n)
    cmd = {"state": pinned((m,), np.uint8), "path_id": pinned((m,), np.int32), "waypoint": pinned((m,), np.int32),
           "r_next": pinned((m, 2), np.float32)}
    idx_t, idx = pinned((m,), np.int32)
    idx_u = idx.view(np.uint32)
    cmd["state"][1][:] = agents["state"][:m]
    cmd["waypoint"][1][:] = 1
    cmd["r_next"][1][:] = np.nan
    cmd_ptrs = {k: v[1].ctypes.data for k, v in cmd.items()}
    h2d = idx.nbytes + sum(v[1].nbytes for v in cmd.values())
    host_board = {k: v[1] for k, v in board.items()}
    chg_t, chg_raw = pinned((n * abi.DELTA_DTYPE.itemsize,), np.uint8)
    changes = chg_raw.view(abi.DELTA_DTYPE)
    moved = []
    rng = np.random.default_rng(7)

    def board_equals_full_download(what):
        full = g.communicate(("vel", "angle_vel", "state", "r_next"))
        for k in full:
            a, b = host_board[k], full[k]
            if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
                raise RuntimeError(f"e2e leg: blackboard field {k} kept through rts_download_delta differs from a full rts_download ({what})")

    def frame(f):
        # the command of this frame: m distinct agents are sent (again) onto the path of their start region
        first = int(rng.integers(0, n - m + 1))
        idx_u[:] = np.arange(first, first + m, dtype=np.uint32)
        cmd["path_id"][1][:] = agents["path_id"][first:first + m]
        cmd["state"][1][:] = agents["state"][first:first + m]
        g.update_shared_data_indexed_ptrs(idx.ctypes.data, m, cmd_ptrs)   # System2::updateSharedData   (H2D)
        g.update(1)                                                      # System2::update
        moved.append(g.communicate_delta(host_board, changes))           # System2::communicate        (D2H, synchronises)

    for f in range(3):
        frame(f)
    board_equals_full_download("before the timed frames")
    Ke = min(K, 100)
    moved.clear()
    t0 = time.perf_counter()
    for f in range(Ke):
        frame(f)
    e2e_s = time.perf_counter() - t0
    board_equals_full_download("after the timed frames")
    per_frame = float(np.mean(moved))
    d2h = int(round(8 * n + 4 + abi.DELTA_DTYPE.itemsize * per_frame))
    return {"value": n * Ke / e2e_s, "unit": "agent-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke,
            "entities_with_changed_fields_per_step": per_frame, "d2h_bytes_per_step_full_download": 21 * n,
            "blackboard_equals_full_download": True,
            "api": "GpuAgentSystem.update_shared_data_indexed -> update -> communicate_delta (rts_update_agents_indexed / rts_step / "
                   "rts_download_delta), pinned host buffers; 1000 commanded agents up; down: vel of every agent + {index, angle_vel, "
                   "target, state} of the entities where one of them changed; the host blackboard {vel, angle_vel, state, target} is "
                   "bit-identical to a full rts_download (checked before and after the timed frames)"}


def config4_single_gpu(args, local, K, Wm):
    """BASELINE config 4 on ONE GPU (the denominator of the 8-GPU scaling figure): 10M agents on the 8192x8192-cell map"""
    from rtsengine_b200 import abi, engine, scenario
    tile = scenario.load_map("map_c3.npz")
    fx = scenario.tile_map(tile, 8, 8)
    n4 = args.agents4
    agents = scenario.tiled_agents(tile, 8, 8, n4 // 64, seed=1237)
    n4 = len(agents["r"])
    p = engine.default_params(8192, 8192)
    p.precision = abi.PRECISION_FAST
    with engine.GpuAgentSystem(p, device=local) as g:
        g.upload_map(scenario.tables_of(fx))
        g.upload_paths(scenario.paths_of(fx))
        ms = []
        for _ in range(3):
            g.set_agents(agents)
            g.update(Wm, sync=True)
            g.reset_timings()
            g.update(K)
            g.sync()
            ms.append(g.timings().last_step_ms)
        med = float(np.median(ms))
    return {"workload": "config4: 10M agents, 8192x8192-cell map (8x8 config-3 tiles, 12 800 buildings, 4096 groups), one GPU, "
                        "tolerance mode", "agents": n4, "value": n4 * K / (med * 1e-3), "unit": "agent-updates/s", "ms_per_step": med / K,
            "repeat_ms_per_step": [m / K for m in ms], "precision": "fast"}


def config5_single_gpu(args, local):
    """BASELINE config 5 on ONE GPU: 4M agents through the corridor map (the crowded-agent kernel and the wall pass at work)"""
    from rtsengine_b200 import abi, engine, scenario
    fx = scenario.load_map("map_c5.npz")
    agents, paths = scenario.config5_agents(fx, n=args.agents5, seed=1238)
    n5 = len(agents["r"])
    p = engine.default_params(2048, 2048)
    p.precision = abi.PRECISION_FAST
    frames = 100
    with engine.GpuAgentSystem(p, device=local) as g:
        g.upload_map(scenario.tables_of(fx))
        g.upload_paths(paths)
        g.set_agents(agents)
        g.update(8, sync=True)
        g.reset_timings()
        g.update(frames)
        g.sync()
        ms = g.timings().last_step_ms
        g.set_agents(agents)
        g.update(8, sync=True)
        g.set_profiling(True)
        g.reset_timings()
        g.update(frames, sync=True)
        kt = {k: g.kernel_time_ms(k)[0] for k in ("step", "dense", "scan", "reorder")}
    return {"workload": f"config5: {n5} agents on the 2048x2048-cell corridor map (64 corridors 6 cells wide), counter-flow, one GPU, tolerance mode",
            "agents": n5, "frames": frames, "value": n5 * frames / (ms * 1e-3), "unit": "agent-updates/s", "ms_per_step": ms / frames,
            "kernel_ms": kt, "precision": "fast"}


def run_gpu(args):
    import torch
    from rtsengine_b200 import abi, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from rtsengine_b200 import slabs
        return slabs.bench_multi_gpu(args, measured_peak, ClockSampler, ALGO_BYTES_PER_UPDATE)

    n = args.agents
    fx, tables, agents, paths = build_world(n)
    K, Wm = args.steps, max(args.warmup, 3)
    peak, peak_src = measured_peak()
    modes = {}
    e2e = None
    with ClockSampler(local) as clocks:
        for prec in (abi.PRECISION_FAST, abi.PRECISION_EXACT):
            params = engine.default_params(1024, 1024)
            params.precision = prec
            with engine.GpuAgentSystem(params, device=local) as g:
                g.upload_map(tables)
                g.upload_paths(paths)
                modes[prec] = time_mode(g, agents, n, K, Wm, args.repeats, PRECISION_NAMES[prec])
                if prec == args.precision:
                    e2e = e2e_leg(g, n, agents, K, Wm)
    head = modes[args.precision]
    other = modes[1 - args.precision]
    roofline = roofline_of(head, n, peak, peak_src, k_step_traffic(n, head["precision"]))
    roofline["other_mode"] = roofline_of(other, n, peak, peak_src, k_step_traffic(n, other["precision"]))
    roofline["other_mode"].update(value=other["value"], ms_per_step=other["ms_per_step"])

    c4 = None
    if not args.no_config4:
        try:
            c4 = config4_single_gpu(args, local, K, Wm)
        except Exception as e:
            c4 = {"value": None, "error": str(e)}

    c5 = None
    if not args.no_config5:
        try:
            c5 = config5_single_gpu(args, local)
        except Exception as e:
            c5 = {"value": None, "error": str(e)}

    cpu = None
    if not args.no_cpu:
        try:
            cpu = cpu_port_baseline(tables, agents, paths, budget_s=args.cpu_budget)
        except Exception as e:  # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "agent-updates/s", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}

    line = {
        "metric": "agent-updates/sec", "value": head["value"], "unit": "agent-updates/s", "n_gpus": 1, "steps": K, "warmup": Wm,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "config3: 1M agents, 64 groups on precomputed funnel paths, 1024x1024-cell CDT map with 200 buildings",
                   "agents": n, "map_cells": [1024, 1024], "buildings": 200, "groups": 64,
                   "precision": head["precision"] + (" (RtsParams.precision = RTS_PRECISION_FAST: reference neighbour sets, wall contacts, waypoint "
                                                     "decisions and index functions; force terms to 1e-5, gated by tests/test_gpu_fast.py)"
                                                     if head["precision"] == "fast" else " (bit-identical to the IEEE build of the reference)"),
                   "timing": f"median of {args.repeats} repetitions of the same {K} frames (initial population + {Wm} warm-up frames each time)",
                   "l2_policy": f"inputs larger than L2: {n} agents x ~190 B touched per frame across ping-pong buffers "
                                f"({n * 190 / 1e6:.0f} MB vs 126 MB L2); no explicit flush"},
        "repeat_ms_per_step": head["repeat_ms_per_step"],
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": head["launches"], "roofline": roofline, "config4_1gpu": c4, "config5_1gpu": c5,
        "cpu_baseline": cpu,
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
        n = 5000
        side = float(np.sqrt(n / (1_000_000 / 5120.0 ** 2)))   # window with config-3 density

        def timed_run(n_threads):
            """a fresh world (same seed, same frames) stepped with n_threads OpenMP threads"""
            L.refh_set_omp_threads(n_threads)
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
            elapsed = time.perf_counter() - t0
            w.close()
            return elapsed

        # the reference hard-codes 6 OpenMP threads for its parallel loops (NUM_OMP_NS_THREADS / NUM_OMP_WALLS_THREADS,
        # src/core.h:27-38); the build here takes the team size from the OpenMP runtime instead, so both are one library:
        # SURVEY §8d asks for the stock setting next to the all-cores one. The line's value is the all-cores run.
        stock = None
        if (os.cpu_count() or 1) >= 6:
            dt6 = timed_run(6)
            stock = {"cores": 6, "value": n * K / dt6, "ms_per_step": 1e3 * dt6 / K,
                     "what": "the same frames with the reference's own hard-coded team size (NUM_OMP_*_THREADS 6, src/core.h:27-38)"}
        dt = timed_run(os.cpu_count() or 1)
        threads = int(refh.load(True).refh_omp_threads())
        kind, sample = "reference", (f"{n} agents (reference cap N_MAX_NAVIGABLE_BOIDS) at config-3 density in a {side:.0f}² window, 50 buildings, "
                                     f"verbatim reference TUs -Ofast, {K} frames")
        value = n * K / dt
        ref_agents = n
        ref_workload = (f"config-3 DENSITY on the reference's own scale: {n} agents (its cap) in a {side:.0f}x{side:.0f}-unit window of a "
                        f"512x512-cell map with 50 buildings, one command group on funnel paths planned by the reference's PathFinder2")
    else:
        from oracle import port
        port.load(True).rts_oracle_set_threads(os.cpu_count() or 1)
        fx, tables, agents, paths = build_world(200_000)
        cpu = cpu_port_baseline(tables, agents, paths, budget_s=20.0)
        value, threads, kind, sample = cpu["value"], cpu["cores"], "port", cpu["sample"]
        dt = None
        ref_agents = 200_000
        ref_workload = "config 3 at 200 000 agents (oracle port; the verbatim reference build is not present on this box)"
    line = {"impl": "reference", "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": n_world,
            "steps": K, "warmup": Wm, "ms_per_step": (1e3 * dt / K) if dt else None, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ref_workload, "agents": ref_agents, "same_config_as_gpu_arm": False,
                       "why": "the reference caps at 5 000 navigable agents (src/core.h:46) and 10 000 entities (src/ECS.h:19): config 3 itself "
                              "(1M agents) cannot run on it; the same-config CPU comparator is the GPU arm's cpu_baseline (oracle port, 1M agents)"},
            "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if kind == "reference" and stock is not None:
        line["cpu_baseline"]["stock_threads"] = stock
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--agents", type=int, default=1_000_000)
    ap.add_argument("--agents4", type=int, default=10_000_000, help="population of the config-4 legs")
    ap.add_argument("--precision", type=int, default=1, choices=[0, 1], help="headline arithmetic: 1 = tolerance mode, 0 = bit-exact mode")
    ap.add_argument("--repeats", type=int, default=5, help="repetitions of the timed window (median reported)")
    ap.add_argument("--no-config4", action="store_true", help="skip the config-4 leg")
    ap.add_argument("--agents5", type=int, default=4_000_000, help="population of the config-5 leg")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
