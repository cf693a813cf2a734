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


def rebalanced_bounds(hist: np.ndarray, bounds: list, cap_migrants: int, max_shift: int = 4, tolerance: float = 0.10) -> list:
    """New slab boundaries from the GLOBAL per-row agent counts: quantile boundaries (as partition_rows), but every
    boundary moves only if its two slabs differ by more than `tolerance`, by at most `max_shift` rows, and by no more rows
    than whose agents fit 80 % of the migrant capacity (they travel through the ordinary migrant buffers in one frame).
    Deterministic: every rank computes the same list from the same all-reduced histogram."""
    hist = np.asarray(hist, np.int64)
    nranks = len(bounds) - 1
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    new = list(bounds)
    for g in range(1, nranks):
        below = int(cum[bounds[g]] - cum[bounds[g - 1]])
        above = int(cum[bounds[g + 1]] - cum[bounds[g]])
        if abs(below - above) <= tolerance * max(below, above, 1):
            continue
        target = int(np.searchsorted(cum, total * g / nranks, side="left"))
        step = int(np.clip(target - bounds[g], -max_shift, max_shift))
        b = bounds[g]
        moved = 0
        while step != 0:
            row = b if step > 0 else b - 1            # the row that changes hands next
            if moved + int(hist[row]) > 0.8 * cap_migrants:
                break
            moved += int(hist[row])
            b += 1 if step > 0 else -1
            step -= 1 if step > 0 else -1
        new[g] = b
    for g in range(1, nranks):                        # keep every slab at least one row tall and the list increasing
        new[g] = min(max(new[g], new[g - 1] + 1), len(hist) - (nranks - g))
    return new


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
        # arrival rule on: the band is finish_radius wide and the halo records carry the navigation record, as in
        # rts_abi.cu::halo_band / ExBuf::h_nav4 (every rank evaluates the rule for its ghosts and stops only agents it owns)
        self.halo_fields = list(HALO_FIELDS)
        if params.enable_arrival:
            self.halo = np.float32(max(float(self.halo), float(params.finish_radius) + 0.25))
            self.halo_fields += ["path_id", "waypoint", "r_next"]
        self.step_fn = step_fn
        self.owned: dict = {}
        self.ghosts = {k: None for k in self.halo_fields}

    def set_agents(self, agents: dict):
        self.owned = {k: np.ascontiguousarray(agents[k]).copy() for k in FULL_FIELDS}
        self._clear_ghosts()

    def _clear_ghosts(self):
        self.ghosts = {"r": np.zeros((0, 2), np.float32), "radius": np.zeros(0, np.float32), "mass": np.zeros(0, np.float32),
                       "state": np.zeros(0, np.uint8), "gid": np.zeros(0, np.uint32), "path_id": np.zeros(0, np.int32),
                       "waypoint": np.zeros(0, np.int32), "r_next": np.zeros((0, 2), np.float32)}
        self.ghosts = {k: self.ghosts[k] for k in self.halo_fields}

    def halo_records(self):
        """(to_down, to_up) halo records of the owned agents' current positions"""
        y = self.owned["r"][:, 1]
        out = []
        for d in (0, 1):
            if not self.has[d]:
                out.append({k: self.owned[k][:0].copy() for k in self.halo_fields})
                continue
            sel = (y < self.y_lo + self.halo) if d == 0 else (y >= self.y_hi - self.halo)
            out.append({k: self.owned[k][sel].copy() for k in self.halo_fields})
        return out

    def add_ghosts(self, rec: dict):
        for k in self.halo_fields:
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
            if k in self.halo_fields:
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
            self.add_ghosts({k: new[k][m].copy() for k in self.halo_fields})
        self.owned = {k: new[k][stay] for k in FULL_FIELDS}
        self.last = {k: v[stay] for k, v in extra.items()}
        halo = self.halo_records()
        return mig[0], mig[1], halo[0], halo[1]

    def set_rows(self, lo: int, hi: int):
        """rts_slab_set_rows: this slab now owns rows [lo, hi). Ownership is decided by frame() from the row of each agent's
        new position, so the agents of the rows that changed hands migrate during the next frame; the halo records of that
        frame are already cut along the new boundary."""
        assert (lo > 0) == self.has[0] and (hi < self.P.ns_ny) == self.has[1], "the outermost slabs keep the map edge"
        self.lo, self.hi = lo, hi
        self.y_lo = np.float32(lo * self.P.ns_cell_size)
        self.y_hi = np.float32(hi * self.P.ns_cell_size)

    def row_counts(self) -> np.ndarray:
        """rts_slab_row_counts: owned agents per physics-grid row"""
        return np.bincount(row_of(self.owned["r"][:, 1], self.P.ns_cell_size, self.P.ns_ny), minlength=int(self.P.ns_ny))

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

    def row_counts(self) -> np.ndarray:
        """owned agents per physics-grid row (whole map: ns_ny entries)"""
        out = np.zeros(int(self.params.ns_ny), np.uint32)
        self.g._chk(self.g.L.rts_slab_row_counts(self.g.h, out.ctypes.data))
        return out

    def set_rows(self, lo: int, hi: int):
        self.g._chk(self.g.L.rts_slab_set_rows(self.g.h, int(lo), int(hi)))
        self.params.slab_row_lo, self.params.slab_row_hi = int(lo), int(hi)

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


ALIGN_FRAMES = 8   # untimed frames between the host barrier and the timed window of a multi-GPU leg (one CUDA-graph replay)


def _slab_case(dist, torch, engine, rank, world, local, fx, agents, n_total, precision, K, Wm, repeats, reserved5=0, clock_sampler=None):
    """One multi-GPU workload: the population cut into `world` y-slabs at agent-count quantiles of physics-grid rows, `repeats`
    repetitions of the same K frames (re-uploaded + Wm warm-up frames each time), device time = max over ranks, median over
    repetitions. Returns (result dict, the SlabRank left open for the caller — close it)."""
    from . import scenario
    tables = scenario.tables_of(fx)
    paths = scenario.paths_of(fx)
    nx, ny = (int(v) for v in fx["n_cells"])
    base = engine.default_params(nx, ny)
    base.precision = precision
    base.reserved[5] = reserved5
    bounds = partition_rows(agents["r"][:, 1], base.ns_ny, base.ns_cell_size, world)
    parts = split_agents(agents, bounds, base.ns_cell_size, base.ns_ny)
    # exchange capacities from the density: agents in a halo band / crossing a boundary per frame, with a wide margin
    per_unit_height = n_total / (ny * 5.0)
    cap_halo = int(max(4096, 6 * per_unit_height * halo_width(base.r_max2)))
    cap_mig = int(max(1024, 6 * per_unit_height * 2.0))
    sr = SlabRank(base, bounds, rank, local, cap_migrants=cap_mig, cap_halo=cap_halo)
    uid = [make_unique_id(sr.g.L) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sr.init_comm(uid[0])
    sr.g.upload_map(tables)
    sr.g.upload_paths(paths)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    ms, launches, wall = [], 0, 0.0
    for rep in range(repeats):
        sr.g.set_agents(parts[rank])
        sr.halo_refresh()
        sr.g.update(Wm, sync=True)
        barrier()                   # every rank enters the timed region together
        # ... on the HOST. The ranks leave the barrier a few hundred microseconds apart (Python), which is 10 % of a 20-frame
        # window; ALIGN_FRAMES further untimed frames are enqueued first, during which the exchange itself brings the
        # devices into step. The K timed frames follow in the same stream without a host synchronisation; their CUDA events
        # bracket exactly K frames on every rank.
        t0 = time.perf_counter()
        sr.g.update(ALIGN_FRAMES)
        sr.g.reset_timings()
        sr.g.update(K)
        sr.g.sync()
        wall = time.perf_counter() - t0
        t = torch.tensor([sr.g.timings().last_step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t.item()))
        launches = int(sr.g.timings().kernel_launches)
    own = sr.owned(("r",))
    cnt = torch.tensor([len(own["gid"])], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt)
    med = float(np.median(ms))
    res = {"value": n_total * K / (med * 1e-3), "unit": "agent-updates/s", "ms_per_step": med / K, "repeat_ms_per_step": [m / K for m in ms],
           "agents": n_total, "slab_rows": bounds, "gpu_launches": launches,
           "conservation": {"agents_after": int(cnt.item()), "agents_before": n_total},
           "exchange": {0: "none", 1: "ncclSend/ncclRecv per frame", 2: "peer memory over NVLink (CUDA IPC: the frame kernels write their records into the neighbour's buffers, one flag word per frame), frames replayed from a CUDA graph"}[
               int(sr.g.L.rts_slab_exchange_mode(sr.g.h))],
           "wall_s_last_repeat": wall, "wall_frames": ALIGN_FRAMES + K}
    return res, sr


def _parity_vs_single_domain(dist, torch, engine, rank, world, local, tile, per_gpu, frames=24):
    """T10 over the REAL inter-GPU path: the slabs (bit-exact mode) against a single-domain handle stepping the same
    population on rank 0; every agent, every float compared as bits. Returns the number of differing agents."""
    from . import scenario
    fx = scenario.tile_map(tile, 1, world)
    n_total = per_gpu * world
    agents = scenario.tiled_agents(tile, 1, world, per_gpu, seed=1236)
    agents["gid"] = np.arange(n_total, dtype=np.uint32)
    res, sr = _slab_case(dist, torch, engine, rank, world, local, fx, agents, n_total, abi.PRECISION_EXACT, frames, 3, 1)
    names = ("r", "inertia", "angle", "angle_vel", "state", "tri_idx", "waypoint", "r_next")
    own = sr.owned(names)
    sr.close()
    dtypes = {name: dt for name, dt, _ in abi.AGENT_FIELDS}
    comps = {name: k for name, _, k in abi.AGENT_FIELDS}
    ref = {}
    if rank == 0:
        base = engine.default_params(*(int(v) for v in fx["n_cells"]))
        with engine.GpuAgentSystem(base, device=local) as one:
            one.upload_map(scenario.tables_of(fx))
            one.upload_paths(scenario.paths_of(fx))
            one.set_agents({k: v for k, v in agents.items() if k != "gid"})
            one.update(3 + ALIGN_FRAMES + frames, sync=True)   # (what _slab_case stepped: warm-up + alignment + timed frames)
            ref = one.communicate(names)
    bad = np.zeros(len(own["gid"]), bool)
    examples = []
    for k in names:
        shape = (n_total, comps[k]) if comps[k] > 1 else (n_total,)
        width = np.dtype(dtypes[k]).itemsize
        t = torch.empty(shape, dtype={4: torch.int32, 1: torch.uint8}[width], device="cuda")
        if rank == 0:
            t.copy_(torch.from_numpy(ref[k].view({4: np.int32, 1: np.uint8}[width]).reshape(shape)))
        dist.broadcast(t, src=0)
        want = t.cpu().numpy()[own["gid"].astype(np.int64)]
        got = own[k].view({4: np.int32, 1: np.uint8}[width]).reshape(want.shape)
        if len(got):
            diff = (got != want).reshape(len(got), -1).any(1)
            bad |= diff
            for i in np.nonzero(diff)[0][:3]:
                examples.append({"rank": rank, "gid": int(own["gid"][i]), "field": k, "slabs": own[k][i].tolist(),
                                 "single_domain": want[i].view(dtypes[k]).tolist(), "r": own["r"][i].tolist()})
    if examples:
        print(json.dumps({"parity_mismatch_examples": examples[:12]}), file=sys.stderr, flush=True)
    n_bad = torch.tensor([int(bad.sum()), len(own["gid"])], device="cuda", dtype=torch.int64)
    dist.all_reduce(n_bad)
    return {"mismatching_agents": int(n_bad[0].item()), "agents_compared": int(n_bad[1].item()), "frames": 3 + ALIGN_FRAMES + frames,
            "fields": list(names), "precision": "exact", "exchange": res["exchange"],
            "what": f"{world} slabs over the real inter-GPU exchange vs ONE single-domain handle on rank 0, same {n_total} agents "
                    f"({world} config-3 tiles), every agent, bitwise"}


def _arrival_parity(dist, torch, engine, rank, world, local, frames=400):
    """The arrival rule (RtsParams.enable_arrival, SeekSystem::finalizeSeek) over the REAL inter-GPU path — peer-memory
    exchange plus the per-frame ncclAllReduce of the groups' stop counts: `world` slabs against ONE single-domain handle on
    rank 0, bit-exact mode, every agent, bitwise. Groups arrive at destinations on the slab boundaries, so agents stop
    same-group neighbours owned by another rank."""
    from . import scenario
    fx = scenario.load_map("map_empty.npz")
    tables = scenario.tables_of(fx)
    nx, ny = (int(v) for v in fx["n_cells"])
    base = engine.default_params(nx, ny)
    base.enable_arrival = 1
    n = 1500 * world
    rng = np.random.default_rng(15)
    a = scenario._blank(n)
    # the map's physics rows cut into `world` equal slabs; one group per boundary (and one in the middle of slab 0)
    rows = int(base.ns_ny)
    bounds = [round(r * rows / world) for r in range(world + 1)]
    H = ny * 5.0
    ends = [[1200.0, b * base.ns_cell_size + rng.uniform(-8, 8)] for b in bounds[1:-1]] + [[1200.0, 0.5 * bounds[1] * base.ns_cell_size]]
    ends = np.array(ends, np.float32)
    g = rng.integers(0, len(ends), n)
    a["r"][:] = ends[g] + np.stack([rng.uniform(-420, -120, n), rng.uniform(-min(180.0, 0.4 * H / world), min(180.0, 0.4 * H / world), n)], 1)
    a["radius"][:] = np.where(rng.random(n) < 0.2, 5.0, 3.0)
    a["state"][:] = abi.MOVING
    a["path_id"][:] = g
    a["r_next"][:] = np.nan
    a["gid"] = np.arange(n, dtype=np.uint32)
    m = len(ends)
    paths = {"offsets": np.arange(m + 1, dtype=np.uint32), "waypoints": ends.copy(), "portals": np.zeros((m, 5), np.float32),
             "path_end": ends.copy(), "group": np.arange(m, dtype=np.int32)}
    parts = split_agents(a, bounds, base.ns_cell_size, base.ns_ny)
    sr = SlabRank(base, bounds, rank, local, cap_migrants=8192, cap_halo=32768)
    uid = [make_unique_id(sr.g.L) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sr.init_comm(uid[0])
    sr.g.upload_map(tables)
    sr.g.upload_paths(paths)
    sr.g.set_agents(parts[rank])
    sr.halo_refresh()
    sr.g.update(frames, sync=True)
    names = ("r", "inertia", "angle", "angle_vel", "state", "path_id", "waypoint", "r_next")
    own = sr.owned(names)
    area = sr.g.finished_area(m)
    mode = int(sr.g.L.rts_slab_exchange_mode(sr.g.h))
    sr.close()
    dtypes = {name: dt for name, dt, _ in abi.AGENT_FIELDS}
    comps = {name: k for name, _, k in abi.AGENT_FIELDS}
    ref, ref_area = {}, np.zeros(m, np.float32)
    if rank == 0:
        with engine.GpuAgentSystem(base, device=local) as one:
            one.upload_map(tables)
            one.upload_paths(paths)
            one.set_agents({k: v for k, v in a.items() if k != "gid"})
            one.update(frames, sync=True)
            ref = one.communicate(names)
            ref_area = one.finished_area(m)
    bad = np.zeros(len(own["gid"]), bool)
    standing = 0
    for k in names:
        shape = (n, comps[k]) if comps[k] > 1 else (n,)
        width = np.dtype(dtypes[k]).itemsize
        t = torch.empty(shape, dtype={4: torch.int32, 1: torch.uint8}[width], device="cuda")
        if rank == 0:
            t.copy_(torch.from_numpy(ref[k].view({4: np.int32, 1: np.uint8}[width]).reshape(shape)))
        dist.broadcast(t, src=0)
        want = t.cpu().numpy()[own["gid"].astype(np.int64)]
        got = own[k].view({4: np.int32, 1: np.uint8}[width]).reshape(want.shape)
        if len(got):
            bad |= (got != want).reshape(len(got), -1).any(1)
        if k == "state":
            standing = int((want == abi.STANDING).sum())
    ta = torch.from_numpy(ref_area.view(np.int32).copy()).cuda()
    dist.broadcast(ta, src=0)
    area_bad = int((ta.cpu().numpy() != area.view(np.int32)).sum())
    tot = torch.tensor([int(bad.sum()), len(own["gid"]), standing, area_bad], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    return {"mismatching_agents": int(tot[0].item()), "agents_compared": int(tot[1].item()), "standing_at_end": int(tot[2].item()),
            "finished_area_mismatches_over_ranks": int(tot[3].item()), "groups": m, "frames": frames, "precision": "exact",
            "exchange_mode": mode,
            "what": f"arrival rule on: {world} slabs (peer exchange + ncclAllReduce of the groups' stop counts every frame) vs ONE "
                    f"single-domain handle on rank 0, {n} agents in {m} groups arriving on the slab boundaries, every agent, bitwise"}


def _config5_case(dist, torch, engine, rank, world, local, n5, precision, frames, chunk, reserved5=0):
    """BASELINE config 5: 4M agents funnelled through the corridor map, `world` slabs, counter-flow; the crowd drains
    unevenly, so every `chunk` frames the per-row agent counts are all-reduced and the slab boundaries follow the quantiles
    (rebalanced_bounds -> rts_slab_set_rows). Device time of the chunks, max over ranks; the re-binning that a boundary move
    costs is reported separately (host wall clock)."""
    from . import scenario
    fx = scenario.load_map("map_c5.npz")
    agents, paths = scenario.config5_agents(fx, n=n5, seed=1238)
    agents["gid"] = np.arange(n5, dtype=np.uint32)
    nx, ny = (int(v) for v in fx["n_cells"])
    base = engine.default_params(nx, ny)
    base.precision = precision
    base.reserved[5] = reserved5
    bounds = partition_rows(agents["r"][:, 1], base.ns_ny, base.ns_cell_size, world)
    parts = split_agents(agents, bounds, base.ns_cell_size, base.ns_ny)
    per_unit_height = n5 / (ny * 5.0)
    # the crowd is far from uniform (a slab boundary can run along a jammed corridor): capacities for 25x the mean density in
    # the halo band, and for one whole physics row of agents at 4x the mean as migrants (rebalancing moves whole rows)
    cap_halo = int(max(65536, 25 * per_unit_height * halo_width(base.r_max2)))
    cap_mig = int(max(65536, 4 * per_unit_height * base.ns_cell_size))
    sr = SlabRank(base, bounds, rank, local, cap_migrants=cap_mig, cap_halo=cap_halo)
    uid = [make_unique_id(sr.g.L) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sr.init_comm(uid[0])
    sr.g.upload_map(scenario.tables_of(fx))
    sr.g.upload_paths(paths)
    sr.g.set_agents(parts[rank])
    sr.halo_refresh()
    sr.g.update(8, sync=True)
    dev_ms, moves, t_reb, imbalance = 0.0, 0, 0.0, []
    done = 0
    while done < frames:
        k = min(chunk, frames - done)
        dist.barrier()
        torch.cuda.synchronize()
        sr.g.reset_timings()
        sr.g.update(k)
        sr.g.sync()
        t = torch.tensor([sr.g.timings().last_step_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms += float(t.item())
        done += k
        if world > 1 and done < frames:
            t0 = time.perf_counter()
            hist = torch.from_numpy(sr.row_counts().astype(np.int64)).cuda()
            dist.all_reduce(hist)
            hist = hist.cpu().numpy()
            cum = np.concatenate([[0], np.cumsum(hist)])
            owned = np.diff(cum[bounds])
            imbalance.append(float(owned.max() / max(owned.mean(), 1)))
            new = rebalanced_bounds(hist, bounds, cap_mig)
            if new != bounds:
                sr.set_rows(new[rank], new[rank + 1])
                bounds = new
                moves += 1
            t_reb += time.perf_counter() - t0
    own = sr.owned(("r",))
    cnt = torch.tensor([len(own["gid"])], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt)
    if os.environ.get("RTS_BENCH_RANK_TRACE"):   # diagnostic: per-rank kernel times of a further (serialised) 20 frames
        sr.g.set_profiling(True)
        sr.g.reset_timings()
        sr.g.update(20, sync=True)
        ms = {k: round(sr.g.kernel_time_ms(k)[0], 4) for k in ("step", "dense", "scan", "reorder", "ingest")}
        print(json.dumps({"config5_rank": rank, "owned": len(own["gid"]), "slots": int(sr.g.n), "kernel_ms": ms}), file=sys.stderr, flush=True)
        sr.g.set_profiling(False)
        dist.barrier()
    res = {"workload": f"config5: {n5} agents on the 2048x2048-cell corridor map (64 corridors 6 cells wide, 1 890 wall blocks), counter-flow, "
                       f"{world} y-slabs re-balanced every {chunk} frames", "agents": n5, "frames": frames,
           "value": n5 * frames / (dev_ms * 1e-3), "unit": "agent-updates/s", "ms_per_step": dev_ms / frames,
           "rebalance": {"every_frames": chunk, "boundary_moves": moves, "host_s_total": t_reb, "max_over_mean_owned": imbalance,
                         "slab_rows_final": bounds},
           "conservation": {"agents_after": int(cnt.item()), "agents_before": n5}, "precision": PRECISION_LABEL[precision]}
    sr.close()
    return res


def _bench_multi_gpu(args, measured_peak, ClockSampler, algo_bytes):
    """N>1 leg of bench.py.
    Headline line (continuity with round 1): WEAK scaling of the N=1 workload — `world` config-3 tiles stacked in y,
    args.agents per GPU. Added: `config4` = BASELINE config 4 itself, STRONG scaling (10M agents on the 8192x8192-cell map
    cut into `world` slabs; bench.py --gpus 1 reports the same population on one GPU as `config4_1gpu`), and
    `parity_vs_single_domain` = the slabs against one domain, bitwise, over the real exchange."""
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
    K, Wm = args.steps, max(args.warmup, 3)
    precision = args.precision
    reserved5 = int(os.environ.get("RTS_RESERVED5", "0"))   # diagnostic A/B of the exchange switches (RtsParams.reserved[5])
    tile = scenario.load_map("map_c3.npz")

    # ---- weak scaling of config 3
    fx = scenario.tile_map(tile, 1, world)
    nx, ny = (int(v) for v in fx["n_cells"])
    agents = scenario.tiled_agents(tile, 1, world, per_gpu, seed=1236)
    agents["gid"] = np.arange(n_total, dtype=np.uint32)
    sampler = ClockSampler(local)   # opens NVML here, outside the timed region (its start-up time differs from rank to rank)
    with sampler as clocks:
        main, sr = _slab_case(dist, torch, engine, rank, world, local, fx, agents, n_total, precision, K, Wm, args.repeats, reserved5)
    if os.environ.get("RTS_BENCH_RANK_TRACE"):   # diagnostic: per-rank kernel times of a further (serialised) 40 frames
        sr.g.set_profiling(True)
        sr.g.reset_timings()
        sr.g.update(40, sync=True)
        ms = {k: round(sr.g.kernel_time_ms(k)[0], 4) for k in engine.KERNELS}
        print(json.dumps({"rank": rank, "kernel_ms": ms}), file=sys.stderr, flush=True)
        sr.g.set_profiling(False)
        dist.barrier()
    sr.close()
    del agents

    # ---- BASELINE config 4, strong scaling
    c4 = None
    if not args.no_config4:
        fx4 = scenario.tile_map(tile, 8, 8)
        a4 = scenario.tiled_agents(tile, 8, 8, args.agents4 // 64, seed=1237)
        n4 = len(a4["r"])
        a4["gid"] = np.arange(n4, dtype=np.uint32)
        c4, sr4 = _slab_case(dist, torch, engine, rank, world, local, fx4, a4, n4, precision, K, Wm, args.repeats, reserved5)
        sr4.close()
        del a4
        c4["workload"] = (f"config4: {n4} agents on the 8192x8192-cell map (8x8 config-3 tiles, 12 800 buildings, 4096 groups) cut into "
                          f"{world} y-slabs; strong scaling — bench.py --gpus 1 reports the same population on one GPU as config4_1gpu")
        c4["precision"] = PRECISION_LABEL[precision]

    # ---- BASELINE config 5: dense corridors, slabs re-balanced while the crowd drains
    c5 = None
    if not args.no_config5:
        c5 = _config5_case(dist, torch, engine, rank, world, local, args.agents5, precision, max(K, 100), int(os.environ.get("RTS_C5_CHUNK", "25")), reserved5)

    # ---- the slabs against a single domain, bitwise, over the real exchange
    parity = None
    if not os.environ.get("RTS_BENCH_NO_PARITY"):
        parity = _parity_vs_single_domain(dist, torch, engine, rank, world, local, tile, min(per_gpu, 250_000),
                                          frames=int(os.environ.get("RTS_PARITY_FRAMES", "24")))

    parity_arrival = None
    if not os.environ.get("RTS_BENCH_NO_PARITY"):
        parity_arrival = _arrival_parity(dist, torch, engine, rank, world, local)

    if rank == 0:
        peak, peak_src = measured_peak()
        line = {
            "metric": "agent-updates/sec", "value": main["value"], "unit": "agent-updates/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"config3 x {world}: {per_gpu} agents per GPU, {world} config-3 tiles stacked in y ({nx}x{ny} cells, "
                                   f"{200 * world} buildings, {64 * world} groups), y-slabs aligned to physics-grid rows, halo + migration "
                                   f"exchange every frame ({main['exchange']})",
                       "agents": n_total, "agents_per_gpu": per_gpu, "slab_rows": main["slab_rows"], "precision": PRECISION_LABEL[precision],
                       "timing": f"median of {args.repeats} repetitions of the same {K} frames (each: re-upload, {Wm} warm-up frames, host barrier, "
                                 f"{ALIGN_FRAMES} untimed frames that bring the devices into step, then the {K} timed frames), max over ranks of the device time",
                       "l2_policy": "inputs larger than L2 (per-GPU state > 126 MB); no explicit flush"},
            "repeat_ms_per_step": main["repeat_ms_per_step"],
            "clocks": clocks.summary(), "gpu_launches": main["gpu_launches"],
            "conservation": main["conservation"],
            "e2e": {"value": n_total * main["wall_frames"] / main["wall_s_last_repeat"], "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "multi-GPU leg keeps state device-resident; host wall-clock around rts_step on rank 0 (no end-to-end claim at N>1)"},
            "roofline": {"bound": "hbm", "kernel": "whole frame", "achieved": algo_bytes * per_gpu / (main["ms_per_step"] * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": algo_bytes * per_gpu / (main["ms_per_step"] * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src},
            "config4": c4, "config5": c5, "parity_vs_single_domain": parity, "parity_arrival_vs_single_domain": parity_arrival,
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


PRECISION_LABEL = {0: "exact", 1: "fast"}
