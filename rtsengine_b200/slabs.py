"""Multi-GPU spatial slab decomposition: host-side logic (one process per GPU / per slab).

The map is cut into y-slabs aligned to physics-grid rows (60 units).  Rank g owns rows
[bounds[g], bounds[g+1]).  Per frame each rank sends to its two neighbours (a) the agents that moved out
of its slab (full record) and (b) the agents within the interaction range of the shared boundary
(neighbour record, "halo"); see include/rts_gpu.h.  The reference has no distributed path
 — this is the north-star extension, and its contract is: results are bit-identical to
the single-slab run, for any number of slabs.

This module holds what is independent of the device: row partitioning, splitting a population, the
protocol rules (shared with the CPU model used by the gloo tests), the NCCL bootstrap through
torch.distributed, and the N>1 leg of bench.py.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

from . import abi

HALO_FACTOR = 1.05
HALO_PAD = 0.25


def halo_width(r_max2: float) -> float:
    """same expression as rts_slab_configure (float32 arithmetic)"""
    return float(np.float32(np.sqrt(np.float32(r_max2)) * np.float32(HALO_FACTOR)) + np.float32(HALO_PAD))


def partition_rows(y: np.ndarray, ns_ny: int, ns_cell: float, nranks: int) -> list:
    """Row boundaries [0=b0 < b1 < ... < b_nranks=ns_ny] with roughly equal agent counts per slab."""
    rows = np.clip(np.floor(np.asarray(y, np.float32) / np.float32(ns_cell)).astype(np.int64), 0, ns_ny - 1)
    hist = np.bincount(rows, minlength=ns_ny)
    cum = np.cumsum(hist)
    total = cum[-1]
    bounds = [0]
    for g in range(1, nranks):
        target = total * g / nranks
        b = int(np.searchsorted(cum, target, side="left")) + 1
        b = max(b, bounds[-1] + 1)
        b = min(b, ns_ny - (nranks - g))
        bounds.append(b)
    bounds.append(ns_ny)
    return bounds


def row_of(y: np.ndarray, ns_cell: float, ns_ny: int) -> np.ndarray:
    """Grid::coordToCell row (floor(y/cs) % ny) for in-map positions"""
    return (np.floor(np.asarray(y, np.float32) / np.float32(ns_cell)).astype(np.int64)) % ns_ny


def split_agents(agents: dict, bounds: list, ns_cell: float, ns_ny: int) -> list:
    """Per-rank agent dicts (with global ids) by the row of each agent's position."""
    n = len(agents["r"])
    gid = agents.get("gid")
    if gid is None:
        gid = np.arange(n, dtype=np.uint32)
    rows = row_of(agents["r"][:, 1], ns_cell, ns_ny)
    out = []
    for g in range(len(bounds) - 1):
        sel = np.nonzero((rows >= bounds[g]) & (rows < bounds[g + 1]))[0]
        part = {k: np.ascontiguousarray(v[sel]) for k, v in agents.items() if k != "gid"}
        part["gid"] = np.ascontiguousarray(gid[sel].astype(np.uint32))
        out.append(part)
    return out


def slab_params(base: abi.RtsParams, lo: int, hi: int) -> abi.RtsParams:
    p = abi.RtsParams()
    C.memmove(C.byref(p), C.byref(base), C.sizeof(abi.RtsParams))
    p.slab_row_lo, p.slab_row_hi = lo, hi
    return p


# ---------------------------------------------------------------------------------------------------------
# Device-independent model of one rank (used with the CPU oracle in tests/test_slabs_gloo.py): the same
# rules as k_step<MULTI> / k_ingest_remote, written with numpy so the protocol can be checked without a GPU.
# ---------------------------------------------------------------------------------------------------------
FULL_FIELDS = ["r", "inertia", "angle", "angle_vel", "radius", "mass", "state", "player", "unit_type", "path_id",
               "waypoint", "r_next", "gid"]
HALO_FIELDS = ["r", "radius", "mass", "state", "gid"]


class SlabModel:
    """One slab stepped by a caller-supplied `step_fn(agents, ghost_mask)` (in-place, caller order = gid order)."""

    def __init__(self, params: abi.RtsParams, lo: int, hi: int, step_fn):
        self.P = params
        self.lo, self.hi = lo, hi
        self.has = (lo > 0, hi < params.ns_ny)
        self.y_lo = np.float32(lo * params.ns_cell_size)
        self.y_hi = np.float32(hi * params.ns_cell_size)
        self.halo = np.float32(halo_width(params.r_max2))
        self.step_fn = step_fn
        self.owned: dict = {}
        self.ghosts = {k: None for k in HALO_FIELDS}

    def set_agents(self, agents: dict):
        self.owned = {k: np.ascontiguousarray(agents[k]).copy() for k in FULL_FIELDS}
        self._clear_ghosts()

    def _clear_ghosts(self):
        self.ghosts = {"r": np.zeros((0, 2), np.float32), "radius": np.zeros(0, np.float32), "mass": np.zeros(0, np.float32),
                       "state": np.zeros(0, np.uint8), "gid": np.zeros(0, np.uint32)}

    def halo_records(self):
        """(to_down, to_up) halo records of the owned agents' current positions"""
        y = self.owned["r"][:, 1]
        out = []
        for d in (0, 1):
            if not self.has[d]:
                out.append({k: self.owned[k][:0].copy() for k in HALO_FIELDS})
                continue
            sel = (y < self.y_lo + self.halo) if d == 0 else (y >= self.y_hi - self.halo)
            out.append({k: self.owned[k][sel].copy() for k in HALO_FIELDS})
        return out

    def add_ghosts(self, rec: dict):
        for k in HALO_FIELDS:
            self.ghosts[k] = np.concatenate([self.ghosts[k], rec[k]])

    def frame(self):
        """one frame over owned + ghosts; returns (migrants_down, migrants_up, halo_down, halo_up)"""
        n_own, n_gh = len(self.owned["gid"]), len(self.ghosts["gid"])
        allg = np.concatenate([self.owned["gid"], self.ghosts["gid"]])
        order = np.argsort(allg, kind="stable")
        names = {f[0] for f in abi.AGENT_FIELDS if f[0] != "gid"}
        a = abi.alloc_agent_arrays(n_own + n_gh, names)
        for k in FULL_FIELDS:
            if k == "gid":
                continue
            if k in HALO_FIELDS:
                full = np.concatenate([self.owned[k], self.ghosts[k]])
            else:
                pad = np.zeros((n_gh,) + self.owned[k].shape[1:], self.owned[k].dtype)
                if k == "path_id":
                    pad[:] = -1
                full = np.concatenate([self.owned[k], pad])
            a[k][:] = full[order]
        ghost_mask = np.concatenate([np.zeros(n_own, np.uint8), np.ones(n_gh, np.uint8)])[order]
        self.step_fn(a, ghost_mask)
        back = np.empty_like(order)
        back[order] = np.arange(len(order))
        own_idx = back[:n_own]
        new = {k: a[k][own_idx].copy() for k in FULL_FIELDS if k != "gid"}
        new["gid"] = self.owned["gid"].copy()
        extra = {k: a[k][own_idx].copy() for k in ("vel", "tri_idx", "flags")}
        rows = row_of(new["r"][:, 1], self.P.ns_cell_size, self.P.ns_ny)
        go_down = (rows < self.lo) & self.has[0]
        go_up = (rows >= self.hi) & self.has[1]
        stay = ~(go_down | go_up)
        mig = [{k: new[k][m].copy() for k in FULL_FIELDS} for m in (go_down, go_up)]
        # emigrants stay one more frame as ghosts on this rank
        self._clear_ghosts()
        for m in (go_down, go_up):
            self.add_ghosts({k: new[k][m].copy() for k in HALO_FIELDS})
        self.owned = {k: new[k][stay] for k in FULL_FIELDS}
        self.last = {k: v[stay] for k, v in extra.items()}
        halo = self.halo_records()
        return mig[0], mig[1], halo[0], halo[1]

    def add_migrants(self, rec: dict):
        for k in FULL_FIELDS:
            self.owned[k] = np.concatenate([self.owned[k], rec[k]])
        if hasattr(self, "last"):
            n = len(rec["gid"])
            self.last["vel"] = np.concatenate([self.last["vel"], np.zeros((n, 2), np.float32)])
            self.last["tri_idx"] = np.concatenate([self.last["tri_idx"], np.full(n, 0xFFFFFFFF, np.uint32)])
            self.last["flags"] = np.concatenate([self.last["flags"], np.zeros(n, np.uint32)])


# ---------------------------------------------------------------------------------------------------------
# GPU ranks
# ---------------------------------------------------------------------------------------------------------
class SlabRank:
    """One GPU slab: GpuAgentSystem configured with slab rows, exchanging through NCCL inside librts_b200."""

    def __init__(self, base_params: abi.RtsParams, bounds: list, rank: int, device: int, cap_migrants=1 << 15,
                 cap_halo=1 << 17):
        from . import engine
        self.engine = engine
        self.rank, self.nranks = rank, len(bounds) - 1
        self.params = slab_params(base_params, bounds[rank], bounds[rank + 1])
        self.g = engine.GpuAgentSystem(self.params, device=device)
        L = self.g.L
        self.g._chk(L.rts_slab_configure(self.g.h, rank, self.nranks, cap_migrants, cap_halo))

    def init_comm(self, unique_id: bytes):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        self.g._chk(self.g.L.rts_comm_init(self.g.h, buf))

    def halo_refresh(self):
        self.g._chk(self.g.L.rts_halo_refresh(self.g.h))

    def owned(self, names=("r",)):
        """download, drop ghosts, sort by gid"""
        names = tuple(dict.fromkeys(tuple(names) + ("gid",)))
        out = self.g.communicate(names)
        keep = out["gid"] != 0xFFFFFFFF
        order = np.argsort(out["gid"][keep], kind="stable")
        return {k: v[keep][order] for k, v in out.items()}

    def close(self):
        self.g.close()


def make_unique_id(engine_lib) -> bytes:
    buf = (C.c_char * 128)()
    rc = engine_lib.rts_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError("rts_comm_unique_id failed (libnccl.so.2 not loadable)")
    return bytes(buf)


def bench_multi_gpu(args, measured_peak, ClockSampler, algo_bytes):
    try:
        return _bench_multi_gpu(args, measured_peak, ClockSampler, algo_bytes)
    except BaseException:
        # a rank that fails must take the job down: its peers are waiting for it inside NCCL and would sit there until
        # the launcher's timeout
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


def _bench_multi_gpu(args, measured_peak, ClockSampler, algo_bytes):
    """N>1 leg of bench.py: weak scaling, per-GPU population fixed (args.agents per GPU, default 1.25M = 10M / 8)."""
    import torch
    import torch.distributed as dist
    from . import engine, scenario

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per_gpu = args.agents
    n_total = per_gpu * world
    # weak scaling of the N=1 workload: `world` config-3 tiles stacked in y (1024 x 1024*world cells), per_gpu agents per
    # tile at config-3 density, 64 command groups per tile; slabs are cut at agent-count quantiles of physics-grid rows
    tile = scenario.load_map("map_c3.npz")
    fx = scenario.tile_map(tile, 1, world)
    tables = scenario.tables_of(fx)
    paths = scenario.paths_of(fx)
    nx, ny = (int(v) for v in fx["n_cells"])
    base = engine.default_params(nx, ny)
    if os.environ.get("RTS_RESERVED5"):   # diagnostic A/B of the stream-overlap switches (RtsParams.reserved[5])
        base.reserved[5] = int(os.environ["RTS_RESERVED5"])
    agents = scenario.tiled_agents(tile, 1, world, per_gpu, seed=1236)
    agents["gid"] = np.arange(n_total, dtype=np.uint32)
    bounds = partition_rows(agents["r"][:, 1], base.ns_ny, base.ns_cell_size, world)
    parts = split_agents(agents, bounds, base.ns_cell_size, base.ns_ny)
    # exchange capacities from the density: agents in a halo band / crossing a boundary per frame, with a wide margin
    per_unit_height = per_gpu / (fx["tile_cells"][1] * 5.0)
    cap_halo = int(max(4096, 6 * per_unit_height * halo_width(base.r_max2)))
    cap_mig = int(max(1024, 6 * per_unit_height * 2.0))
    sr = SlabRank(base, bounds, rank, local, cap_migrants=cap_mig, cap_halo=cap_halo)
    uid = [make_unique_id(sr.g.L) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sr.init_comm(uid[0])
    sr.g.upload_map(tables)
    sr.g.upload_paths(paths)
    sr.g.set_agents(parts[rank])
    sr.halo_refresh()
    K, Wm = args.steps, max(args.warmup, 3)
    sr.g.update(Wm, sync=True)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)   # opens NVML here, outside the timed region (its start-up time differs from rank to rank)
    with sampler as clocks:
        barrier()                   # every rank enters the timed region together
        sr.g.reset_timings()
        t0 = time.perf_counter()
        sr.g.update(K)
        sr.g.sync()
        wall = time.perf_counter() - t0
    dev_ms = sr.g.timings().last_step_ms
    launches = int(sr.g.timings().kernel_launches)
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    own = sr.owned(("r",))
    cnt = torch.tensor([len(own["gid"])], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt)
    barrier()
    if os.environ.get("RTS_BENCH_RANK_TRACE"):   # diagnostic: per-rank kernel times of a further (serialised) 40 frames
        sr.g.set_profiling(True)
        sr.g.reset_timings()
        sr.g.update(40, sync=True)
        ms = {k: round(sr.g.kernel_time_ms(k)[0], 4) for k in engine.KERNELS}
        print(json.dumps({"rank": rank, "dev_ms_per_frame": dev_ms / K, "owned": len(own["gid"]), "kernel_ms": ms}),
              file=sys.stderr, flush=True)
        sr.g.set_profiling(False)
        barrier()
    if rank == 0:
        peak, peak_src = measured_peak()
        value = n_total * K / (max_ms * 1e-3)
        line = {
            "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"config3 x {world}: {per_gpu} agents per GPU, {world} config-3 tiles stacked in y ({nx}x{ny} cells, "
                                   f"{200 * world} buildings, {64 * world} groups), y-slabs aligned to physics-grid rows, NCCL halo + migration "
                                   f"exchange every frame",
                       "agents": n_total, "agents_per_gpu": per_gpu, "slab_rows": bounds,
                       "l2_policy": "inputs larger than L2 (per-GPU state > 126 MB); no explicit flush"},
            "clocks": clocks.summary(), "gpu_launches": launches,
            "conservation": {"agents_after": int(cnt.item()), "agents_before": n_total},
            "e2e": {"value": n_total * K / wall, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "multi-GPU leg keeps state device-resident; host wall-clock around rts_step on rank 0"},
            "roofline": {"bound": "hbm", "kernel": "whole frame", "achieved": algo_bytes * per_gpu / (max_ms / K * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": algo_bytes * per_gpu / (max_ms / K * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    sr.close()
    dist.destroy_process_group()
