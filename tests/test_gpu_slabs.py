"""GPU (one device): two slab handles emulate two ranks on one GPU — the buffers that NCCL would carry are copied
device-to-device by rts_slab_exchange_local — and must reproduce the single-handle run bit for bit."""
import numpy as np
import pytest

from rtsengine_b200 import abi, engine, slabs
from tests import helpers as H
from tests.test_gpu_parity import random_world

pytestmark = pytest.mark.gpu


# (the third case is a crowd of 1.4 agents per unit² cut by the slab boundary: the edge rows' own queue of crowded agents)
# split: the opt-in frame layout that steps the edge rows first on a comm stream (RtsParams.reserved[5] bit 3)
@pytest.mark.parametrize("split", [0, 8])
@pytest.mark.parametrize("n,lo,hi,frames", [(40000, 100.0, 1100.0, 12), (6000, 380.0, 620.0, 8), (20000, 440.0, 560.0, 6)])
def test_two_slabs_equal_one_domain(n, lo, hi, frames, split):
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    a, paths = random_world(21, n, tables, lo, hi)
    a["inertia"][:] *= 4                      # faster agents: more migration across the cut
    base = engine.default_params(512, 512)
    names = ("r", "inertia", "angle", "angle_vel", "state", "tri_idx", "waypoint", "r_next", "flags")
    with engine.GpuAgentSystem(base) as one:
        one.upload_map(tables)
        one.upload_paths(paths)
        one.set_agents(a)
        bounds = slabs.partition_rows(a["r"][:, 1], base.ns_ny, base.ns_cell_size, 2)
        a2 = dict(a)
        a2["gid"] = np.arange(n, dtype=np.uint32)
        parts = slabs.split_agents(a2, bounds, base.ns_cell_size, base.ns_ny)
        slab_base = engine.default_params(512, 512)
        slab_base.reserved[5] = split
        ranks = [slabs.SlabRank(slab_base, bounds, r, 0, cap_migrants=4096, cap_halo=16384) for r in range(2)]
        L = ranks[0].g.L
        try:
            for r, sr in enumerate(ranks):
                sr.g.upload_map(tables)
                sr.g.upload_paths(paths)
                sr.g.set_agents(parts[r])
            # initial halo
            for sr in ranks:
                sr.g._chk(L.rts_halo_begin(sr.g.h))
            ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
            for sr in ranks:
                sr.g._chk(L.rts_step_end(sr.g.h))
            migrated = 0
            for f in range(frames):
                one.update(1)
                for sr in ranks:
                    sr.g._chk(L.rts_step_begin(sr.g.h))
                ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
                for sr in ranks:
                    sr.g._chk(L.rts_step_end(sr.g.h))
                ref = one.communicate(names)
                got = [sr.owned(names) for sr in ranks]
                gid = np.concatenate([g["gid"] for g in got])
                assert len(gid) == n and len(np.unique(gid)) == n, "agents are conserved, each owned exactly once"
                order = np.argsort(gid)
                finite = np.isfinite(ref["r"]).all(1)
                for k in names:
                    merged = np.concatenate([g[k] for g in got])[order]
                    H.assert_same_bits(merged[finite], ref[k][finite], f"frame {f} {k}")
                migrated += abs(len(got[0]["gid"]) - len(parts[0]["gid"]))
                # ownership follows the physics-grid row of the position
                for r, g in enumerate(got):
                    rows = slabs.row_of(g["r"][:, 1], base.ns_cell_size, base.ns_ny)
                    ok = np.isfinite(g["r"][:, 1])
                    assert ((rows[ok] >= bounds[r]) & (rows[ok] < bounds[r + 1])).all()
        finally:
            for sr in ranks:
                sr.close()


def test_exchange_capacity_overflow_is_reported():
    """fixed-capacity halo / migrant buffers: an overflow must surface as RTS_E_CAPACITY, never as silent loss"""
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    n = 20000
    a, paths = random_world(22, n, tables, 100.0, 1100.0)
    base = engine.default_params(512, 512)
    bounds = slabs.partition_rows(a["r"][:, 1], base.ns_ny, base.ns_cell_size, 2)
    a2 = dict(a)
    a2["gid"] = np.arange(n, dtype=np.uint32)
    parts = slabs.split_agents(a2, bounds, base.ns_cell_size, base.ns_ny)
    ranks = [slabs.SlabRank(base, bounds, r, 0, cap_migrants=4, cap_halo=8) for r in range(2)]   # far too small
    L = ranks[0].g.L
    try:
        for r, sr in enumerate(ranks):
            sr.g.upload_map(tables)
            sr.g.upload_paths(paths)
            sr.g.set_agents(parts[r])
        for sr in ranks:
            sr.g._chk(L.rts_halo_begin(sr.g.h))
        with pytest.raises(engine.RtsError) as e:
            for sr in ranks:
                sr.g.sync()
        assert e.value.code == abi.RTS_E_CAPACITY
    finally:
        for sr in ranks:
            sr.close()


def test_slab_handle_requires_global_ids():
    fx = H.load("frame_small.npz")
    base = engine.default_params(512, 512)
    sr = slabs.SlabRank(base, [0, 20, 43], 0, 0, cap_migrants=64, cap_halo=64)
    try:
        with pytest.raises(engine.RtsError) as e:
            sr.g.set_agents({"r": np.ones((10, 2), np.float32)})
        assert e.value.code == abi.RTS_E_ARG
    finally:
        sr.close()


def test_agents_outside_the_map_do_not_break_the_edge_row_split():
    """Positions beyond the map are clamped into the first / last fine row of a slab's grid and their physics row wraps
    (Grid.cpp:18-24), so they can pack from rows far from the slab boundary: those rows always belong to the edge launch.
    The frame must neither lose an agent nor report RTS_E_STATE; NaN agents stay with their owner."""
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    n = 6000
    a, paths = random_world(23, n, tables, 200.0, 2300.0)
    a["r"][:40, 1] = np.linspace(-30.0, -1.0, 40, dtype=np.float32)          # below the map (slab 0, no lower neighbour)
    a["r"][40:80, 1] = np.linspace(2561.0, 2600.0, 40, dtype=np.float32)     # above it (slab 1, no upper neighbour)
    a["r"][80:84] = a["r"][84:88]                                            # coincident pairs: NaN after one frame
    base = engine.default_params(512, 512)
    base.reserved[5] = 8                                                     # the split frame layout
    bounds = slabs.partition_rows(a["r"][:, 1], base.ns_ny, base.ns_cell_size, 2)
    a2 = dict(a)
    a2["gid"] = np.arange(n, dtype=np.uint32)
    parts = slabs.split_agents(a2, bounds, base.ns_cell_size, base.ns_ny)
    ranks = [slabs.SlabRank(base, bounds, r, 0, cap_migrants=4096, cap_halo=16384) for r in range(2)]
    L = ranks[0].g.L
    try:
        for r, sr in enumerate(ranks):
            sr.g.upload_map(tables)
            sr.g.upload_paths(paths)
            sr.g.set_agents(parts[r])
        for sr in ranks:
            sr.g._chk(L.rts_halo_begin(sr.g.h))
        ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
        for sr in ranks:
            sr.g._chk(L.rts_step_end(sr.g.h))
        for f in range(10):
            for sr in ranks:
                sr.g._chk(L.rts_step_begin(sr.g.h))
            ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
            for sr in ranks:
                sr.g._chk(L.rts_step_end(sr.g.h))
            got = [sr.owned(("r",)) for sr in ranks]           # synchronises: raises on RTS_E_STATE / RTS_E_CAPACITY
            gid = np.concatenate([g["gid"] for g in got])
            assert len(gid) == n and len(np.unique(gid)) == n, f"frame {f}: agents are conserved"
    finally:
        for sr in ranks:
            sr.close()


def test_moving_the_slab_boundary_keeps_the_result_bit_identical():
    """Rebalancing (rts_slab_row_counts / rts_slab_set_rows): the boundary between two slabs is moved twice during the run;
    the agents of the rows that change hands migrate through the ordinary buffers, and every frame still equals the
    single-domain run bit for bit. The row histogram of the two handles adds up to the population."""
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    n = 30000
    a, paths = random_world(23, n, tables, 100.0, 1100.0)
    base = engine.default_params(512, 512)
    names = ("r", "inertia", "angle", "angle_vel", "state", "tri_idx", "waypoint", "r_next", "flags")
    with engine.GpuAgentSystem(base) as one:
        one.upload_map(tables)
        one.upload_paths(paths)
        one.set_agents(a)
        bounds = slabs.partition_rows(a["r"][:, 1], base.ns_ny, base.ns_cell_size, 2)
        a2 = dict(a)
        a2["gid"] = np.arange(n, dtype=np.uint32)
        parts = slabs.split_agents(a2, bounds, base.ns_cell_size, base.ns_ny)
        ranks = [slabs.SlabRank(engine.default_params(512, 512), bounds, r, 0, cap_migrants=8192, cap_halo=16384) for r in range(2)]
        L = ranks[0].g.L
        try:
            for r, sr in enumerate(ranks):
                sr.g.upload_map(tables)
                sr.g.upload_paths(paths)
                sr.g.set_agents(parts[r])
            for sr in ranks:
                sr.g._chk(L.rts_halo_begin(sr.g.h))
            ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
            for sr in ranks:
                sr.g._chk(L.rts_step_end(sr.g.h))
            moves = {3: +2, 7: -3}                      # frame -> rows the boundary moves by
            for f in range(12):
                if f in moves:
                    for sr in ranks:
                        sr.g.sync()
                    hist = ranks[0].row_counts().astype(np.int64) + ranks[1].row_counts().astype(np.int64)
                    assert hist.sum() == n
                    bounds = [bounds[0], bounds[1] + moves[f], bounds[2]]
                    changing = hist[min(bounds[1], bounds[1] - moves[f]):max(bounds[1], bounds[1] - moves[f])].sum()
                    assert 0 < changing < 8192
                    ranks[0].set_rows(bounds[0], bounds[1])
                    ranks[1].set_rows(bounds[1], bounds[2])
                one.update(1)
                for sr in ranks:
                    sr.g._chk(L.rts_step_begin(sr.g.h))
                ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
                for sr in ranks:
                    sr.g._chk(L.rts_step_end(sr.g.h))
                ref = one.communicate(names)
                got = [sr.owned(names) for sr in ranks]
                gid = np.concatenate([g["gid"] for g in got])
                assert len(gid) == n and len(np.unique(gid)) == n, "agents are conserved, each owned exactly once"
                order = np.argsort(gid)
                finite = np.isfinite(ref["r"]).all(1)
                for k in names:
                    merged = np.concatenate([g[k] for g in got])[order]
                    H.assert_same_bits(merged[finite], ref[k][finite], f"frame {f} {k}")
                for r, g in enumerate(got):
                    rows = slabs.row_of(g["r"][:, 1], base.ns_cell_size, base.ns_ny)
                    ok = np.isfinite(g["r"][:, 1])
                    if f - 1 not in moves and f not in moves:      # (one frame after a move the hand-over is complete)
                        assert ((rows[ok] >= bounds[r]) & (rows[ok] < bounds[r + 1])).all()
        finally:
            for sr in ranks:
                sr.close()


def test_arrival_rule_on_two_slabs_equals_one_domain():
    """SeekSystem::finalizeSeek across a slab boundary (the group's finished area is global, CommandGroupManager.h:
    163-172): two groups arrive at destinations ON the cut, so arriving agents stop same-group neighbours owned by the other
    slab, and both slabs must see the same finished area every frame. Two slabs (local exchange emulation, which also sums
    the per-group stop counts as ncclAllReduce does between ranks) == one domain, every agent, bitwise, over the whole arrival."""
    fx = H.load("map_empty.npz")
    tables = H.tables_of(fx)
    n = 3000
    rng = np.random.default_rng(15)
    a = abi.alloc_agent_arrays(n, set(H.ALL_FIELDS))
    a["r"][:] = np.stack([rng.uniform(780, 1080, n), rng.uniform(720, 1080, n)], 1)
    a["radius"][:] = np.where(rng.random(n) < 0.2, 5.0, 3.0)
    a["mass"][:] = 1
    a["state"][:] = abi.MOVING
    a["path_id"][:] = rng.integers(0, 2, n)
    a["r_next"][:] = np.nan
    base = engine.default_params(512, 512)
    cut_row = 15                                            # y = 900
    ends = np.array([[1150.0, 900.0], [1230.0, 893.0]], np.float32)   # both blobs straddle the cut
    paths = {"offsets": np.array([0, 1, 2], np.uint32), "waypoints": ends.copy(), "portals": np.zeros((2, 5), np.float32),
             "path_end": ends.copy(), "group": np.array([0, 1], np.int32)}
    base.enable_arrival = 1
    names = ("r", "inertia", "angle", "angle_vel", "state", "path_id", "waypoint", "r_next", "flags")
    bounds = [0, cut_row, int(base.ns_ny)]
    with engine.GpuAgentSystem(base) as one:
        one.upload_map(tables)
        one.upload_paths(paths)
        one.set_agents(a)
        a2 = dict(a)
        a2["gid"] = np.arange(n, dtype=np.uint32)
        parts = slabs.split_agents(a2, bounds, base.ns_cell_size, base.ns_ny)
        assert min(len(p["gid"]) for p in parts) > 500
        ranks = [slabs.SlabRank(base, bounds, r, 0, cap_migrants=4096, cap_halo=16384) for r in range(2)]
        L = ranks[0].g.L
        try:
            for r, sr in enumerate(ranks):
                sr.g.upload_map(tables)
                sr.g.upload_paths(paths)
                sr.g.set_agents(parts[r])
            for sr in ranks:
                sr.g._chk(L.rts_halo_begin(sr.g.h))
            ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
            for sr in ranks:
                sr.g._chk(L.rts_step_end(sr.g.h))
            standing, cross = [], 0
            for f in range(600):
                one.update(1)
                for sr in ranks:
                    sr.g._chk(L.rts_step_begin(sr.g.h))
                ranks[0].g._chk(L.rts_slab_exchange_local(ranks[0].g.h, ranks[1].g.h))
                for sr in ranks:
                    sr.g._chk(L.rts_step_end(sr.g.h))
                if f % 20 == 19:
                    ref = one.communicate(names)
                    got = [sr.owned(names) for sr in ranks]
                    gid = np.concatenate([g["gid"] for g in got])
                    assert len(gid) == n and len(np.unique(gid)) == n
                    order = np.argsort(gid)
                    for k in names:
                        H.assert_same_bits(np.concatenate([g[k] for g in got])[order], ref[k], f"frame {f} {k}")
                    area = one.finished_area(2)
                    for sr in ranks:
                        H.assert_same_bits(sr.g.finished_area(2), area, f"frame {f}: finished area per group on every slab")
                    standing.append(int((ref["state"] == abi.STANDING).sum()))
                    cross = max(cross, min(int((g["state"] == abi.STANDING).sum()) for g in got))
            assert standing[0] == 0 and standing[-1] > 0.8 * n, standing
            assert cross > 100, "agents stopped on both sides of the cut"
        finally:
            for sr in ranks:
                sr.close()
