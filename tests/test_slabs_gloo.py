"""CPU, world_size 2 over gloo: the slab protocol (ownership by physics-grid row, halo band, one-frame ghosts of
emigrants, migration) modelled with numpy around the CPU oracle must reproduce the single-domain oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _send(rec, dst, fields):
    hdr = torch.tensor([len(rec["gid"])], dtype=torch.int64)
    dist.send(hdr, dst)
    for k in fields:
        t = torch.from_numpy(np.ascontiguousarray(rec[k]).view(np.uint8).reshape(-1).copy())
        if t.numel():
            dist.send(t, dst)


def _recv(src, fields, template):
    hdr = torch.zeros(1, dtype=torch.int64)
    dist.recv(hdr, src)
    m = int(hdr.item())
    out = {}
    for k in fields:
        shape = (m,) + template[k].shape[1:]
        buf = np.zeros(shape, template[k].dtype)
        if m:
            t = torch.zeros(buf.nbytes, dtype=torch.uint8)
            dist.recv(t, src)
            buf = t.numpy().view(template[k].dtype).reshape(shape).copy()
        out[k] = buf
    return out


def _worker(rank, world, port_no, frames, result_dir, start_bounds=None, rebalance_every=0, cap_migrants=0):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port
    from rtsengine_b200 import slabs
    from tests.test_gpu_parity import random_world

    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    n = 6000
    a, paths = random_world(33, n, tables, 150.0, 750.0)
    a["inertia"][:] *= 4
    P = port.default_params(512, 512)
    ow = port.OracleWorld(P, tables, paths)
    a["gid"] = np.arange(n, dtype=np.uint32)
    bounds = list(start_bounds) if start_bounds else slabs.partition_rows(a["r"][:, 1], P.ns_ny, P.ns_cell_size, world)
    parts = slabs.split_agents(a, bounds, P.ns_cell_size, P.ns_ny)
    m = slabs.SlabModel(P, bounds[rank], bounds[rank + 1], lambda ag, gh: ow.step(ag, ghost=gh))
    history = [list(bounds)]
    m.set_agents(parts[rank])
    nb = {0: rank - 1, 1: rank + 1}

    def exchange(out_by_dir, fields, template):
        got = {}
        for d in (0, 1):                      # even ranks send first, odd ranks receive first: no deadlock over gloo p2p
            if not m.has[d]:
                continue
            if rank % 2 == 0:
                _send(out_by_dir[d], nb[d], fields)
                got[d] = _recv(nb[d], fields, template)
            else:
                got[d] = _recv(nb[d], fields, template)
                _send(out_by_dir[d], nb[d], fields)
        return got

    halo = m.halo_records()
    for rec in exchange({0: halo[0], 1: halo[1]}, slabs.HALO_FIELDS, parts[rank]).values():
        m.add_ghosts(rec)
    for f in range(frames):
        md, mu, hd, hu = m.frame()
        for rec in exchange({0: md, 1: mu}, slabs.FULL_FIELDS, parts[rank]).values():
            m.add_migrants(rec)
        for rec in exchange({0: hd, 1: hu}, slabs.HALO_FIELDS, parts[rank]).values():
            m.add_ghosts(rec)
        if rebalance_every and (f + 1) % rebalance_every == 0 and f + 1 < frames:
            # the loop of slabs.run_config5: global row histogram (all-reduce) -> rebalanced_bounds -> rts_slab_set_rows
            hist = torch.from_numpy(m.row_counts().astype(np.int64))
            dist.all_reduce(hist)
            new = slabs.rebalanced_bounds(hist.numpy(), bounds, cap_migrants)
            if new != bounds:
                m.set_rows(new[rank], new[rank + 1])
                bounds = new
                history.append(list(bounds))
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), bounds_history=np.array(history), **m.owned)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])          # 3: the middle rank has a neighbour on either side
def test_slab_protocol_world_size_2(tmp_path, world):
    frames = 10
    port_no = 29500 + 2000 * (world - 2) + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port_no, frames, str(tmp_path)), nprocs=world, join=True)
    # single-domain oracle
    from oracle import port
    from tests.test_gpu_parity import random_world
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    a, paths = random_world(33, 6000, tables, 150.0, 750.0)
    a["inertia"][:] *= 4
    ow = port.OracleWorld(port.default_params(512, 512), tables, paths)
    r0 = a["r"].copy()
    for _ in range(frames):
        ow.step(a)
    parts = [dict(np.load(os.path.join(tmp_path, f"rank{r}.npz"))) for r in range(world)]
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == 6000 and len(np.unique(gid)) == 6000
    order = np.argsort(gid)
    for k in ("r", "inertia", "angle", "angle_vel", "state", "waypoint", "r_next"):
        merged = np.concatenate([p[k] for p in parts])[order]
        H.assert_same_bits(merged, a[k], k)
    # every cut was actually crossed
    from rtsengine_b200 import slabs
    P = port.default_params(512, 512)
    bounds = slabs.partition_rows(r0[:, 1], P.ns_ny, P.ns_cell_size, world)
    start_rows = slabs.row_of(r0[:, 1], P.ns_cell_size, P.ns_ny)
    end_rows = slabs.row_of(a["r"][:, 1], P.ns_cell_size, P.ns_ny)
    for b in bounds[1:-1]:
        assert ((start_rows < b) != (end_rows < b)).sum() > 0


# ---------------------------------------------------------------------------------------------------------
# The arrival rule across a slab boundary (SeekSystem::finalizeSeek, SeekSystem.cpp:235-275; the group's finished
# area is global, CommandGroupManager.h:163-172): the rules of k_arrive<multi> / k_area_sync_counts / the per-frame
# all-reduce of the stop counts, modelled with the oracle (ghosts evaluate the rule, only owners stop and count) and
# dist.all_reduce over gloo.
# ---------------------------------------------------------------------------------------------------------
ARRIVAL_N, ARRIVAL_FRAMES, ARRIVAL_CUT = 1500, 420, 15


def _arrival_world():
    from rtsengine_b200 import abi
    from oracle import port
    fx = H.load("map_empty.npz")
    tables = H.tables_of(fx)
    n = ARRIVAL_N
    rng = np.random.default_rng(15)
    a = abi.alloc_agent_arrays(n, set(H.ALL_FIELDS))
    a["r"][:] = np.stack([rng.uniform(900, 1080, n), rng.uniform(780, 1020, n)], 1)
    a["radius"][:] = np.where(rng.random(n) < 0.2, 5.0, 3.0)
    a["mass"][:] = 1
    a["state"][:] = abi.MOVING
    a["path_id"][:] = rng.integers(0, 2, n)
    a["r_next"][:] = np.nan
    ends = np.array([[1150.0, 900.0], [1230.0, 893.0]], np.float32)     # both destinations lie on the cut (y = 900)
    paths = {"offsets": np.array([0, 1, 2], np.uint32), "waypoints": ends.copy(), "portals": np.zeros((2, 5), np.float32),
             "path_end": ends.copy(), "group": np.array([0, 1], np.int32)}
    P = port.default_params(512, 512)
    P.enable_arrival = 1
    return P, tables, paths, a


def _arrival_worker(rank, world, port_no, frames, result_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port
    from rtsengine_b200 import slabs

    P, tables, paths, a = _arrival_world()
    ow = port.OracleWorld(P, tables, paths)
    counts = np.zeros(2, np.uint32)
    ow.set_slab_arrival(True, counts)
    a["gid"] = np.arange(len(a["r"]), dtype=np.uint32)
    bounds = [0, ARRIVAL_CUT, int(P.ns_ny)]
    parts = slabs.split_agents(a, bounds, P.ns_cell_size, P.ns_ny)
    m = slabs.SlabModel(P, bounds[rank], bounds[rank + 1], lambda ag, gh: ow.step(ag, ghost=gh))
    assert m.halo >= P.finish_radius and "path_id" in m.halo_fields
    m.set_agents(parts[rank])
    other = 1 - rank

    def exchange(rec, fields):
        if rank == 0:
            _send(rec, other, fields)
            return _recv(other, fields, parts[rank])
        got = _recv(other, fields, parts[rank])
        _send(rec, other, fields)
        return got

    d = 1 if rank == 0 else 0                 # direction of the one neighbour
    m.add_ghosts(exchange(m.halo_records()[d], m.halo_fields))
    area_unit = np.float32(np.float32(np.pi) * np.float32(P.rhard)) * np.float32(P.rhard)
    area = np.zeros(2, np.float32)            # k_area_sync_counts: one addition of the unit per stop, in any order
    applied = np.zeros(2, np.int64)
    areas = []
    for f in range(frames):
        out = m.frame()
        t = torch.from_numpy(counts.astype(np.int64))
        dist.all_reduce(t)                    # ncclAllReduce of the stop counts (rts_abi.cu, "the one collective on the frame path")
        total = t.numpy()
        for g in range(2):
            for _ in range(int(total[g] - applied[g])):
                area[g] = area[g] + area_unit
        applied[:] = total
        ow.fin_area[:] = area                 # every rank enters the next frame with the global finished area
        areas.append(area.copy())
        m.add_migrants(exchange(out[d], slabs.FULL_FIELDS))
        m.add_ghosts(exchange(out[2 + d], m.halo_fields))
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), areas=np.array(areas), own_stops=counts, **m.owned)
    dist.destroy_process_group()


def test_arrival_rule_across_the_cut_world_size_2(tmp_path):
    from rtsengine_b200 import abi
    from oracle import port
    frames = ARRIVAL_FRAMES
    port_no = 35500 + (os.getpid() % 2000)
    mp.spawn(_arrival_worker, args=(2, port_no, frames, str(tmp_path)), nprocs=2, join=True)
    P, tables, paths, a = _arrival_world()
    ow = port.OracleWorld(P, tables, paths)
    ref_areas = []
    for _ in range(frames):
        ow.step(a)
        ref_areas.append(ow.fin_area.copy())
    parts = [dict(np.load(os.path.join(tmp_path, f"rank{r}.npz"))) for r in range(2)]
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == ARRIVAL_N and len(np.unique(gid)) == ARRIVAL_N
    order = np.argsort(gid)
    for k in ("r", "inertia", "angle", "angle_vel", "state", "path_id", "waypoint", "r_next"):
        H.assert_same_bits(np.concatenate([p[k] for p in parts])[order], a[k], k)
    for p in parts:                           # the same finished area on every rank, every frame
        H.assert_same_bits(p["areas"], np.array(ref_areas), "finished area per frame")
    standing = int((a["state"] == abi.STANDING).sum())
    assert standing > 0.3 * ARRIVAL_N, standing
    assert min(int(p["own_stops"].sum()) for p in parts) > 50, "agents were stopped on both sides of the cut"
    assert sum(int(p["own_stops"].sum()) for p in parts) == standing   # each agent stopped, and counted, exactly once


def test_slab_rebalancing_world_size_2(tmp_path):
    """Rows change hands at run time (boundaries follow the agent-count quantiles every K frames): a lopsided
    start (rank 0 owns 3 of 43 rows), the histogram all-reduced over gloo, rebalanced_bounds on every rank, the moved rows
    migrating through the ordinary migrant records — and still every agent bit-identical to the single-domain oracle."""
    from oracle import port
    from rtsengine_b200 import slabs
    from tests.test_gpu_parity import random_world
    frames = 12
    P = port.default_params(512, 512)
    start = [0, 3, int(P.ns_ny)]
    port_no = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port_no, frames, str(tmp_path), start, 3, 4000), nprocs=2, join=True)
    fx = H.load("frame_small.npz")
    tables = H.tables_of(fx)
    a, paths = random_world(33, 6000, tables, 150.0, 750.0)
    a["inertia"][:] *= 4
    ow = port.OracleWorld(P, tables, paths)
    first_rows = slabs.row_of(a["r"][:, 1], P.ns_cell_size, P.ns_ny)
    for _ in range(frames):
        ow.step(a)
    parts = [dict(np.load(os.path.join(tmp_path, f"rank{r}.npz"))) for r in range(2)]
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == 6000 and len(np.unique(gid)) == 6000
    order = np.argsort(gid)
    for k in ("r", "inertia", "angle", "angle_vel", "state", "waypoint", "r_next"):
        H.assert_same_bits(np.concatenate([p[k] for p in parts])[order], a[k], k)
    hist = [h.tolist() for h in parts[0]["bounds_history"]]
    assert hist == [h.tolist() for h in parts[1]["bounds_history"]], "both ranks took the same decisions"
    assert len(hist) >= 3 and all(h2[1] > h1[1] for h1, h2 in zip(hist, hist[1:])), hist   # the cut climbed towards the median row
    assert all(abs(h2[1] - h1[1]) <= 4 for h1, h2 in zip(hist, hist[1:]))
    before = int((first_rows < start[1]).sum())
    after = len(parts[0]["gid"])
    assert before < 0.2 * 6000 and after > 0.4 * 6000, (before, after)
    end_rows = slabs.row_of(a["r"][:, 1], P.ns_cell_size, P.ns_ny)
    final = hist[-1]
    # the last boundary move is at least two frames old: every agent sits on its owner again
    for r, p in enumerate(parts):
        rows = slabs.row_of(p["r"][:, 1], P.ns_cell_size, P.ns_ny)
        assert ((rows >= final[r]) & (rows < final[r + 1])).all()
    assert (end_rows < final[1]).sum() == after


def test_partition_and_split():
    from rtsengine_b200 import slabs
    rng = np.random.default_rng(0)
    y = rng.uniform(0, 2560, 10000).astype(np.float32)
    b = slabs.partition_rows(y, 43, 60.0, 4)
    assert b[0] == 0 and b[-1] == 43 and all(b[i] < b[i + 1] for i in range(4))
    agents = {"r": np.stack([y, y], 1)}
    parts = slabs.split_agents(agents, b, 60.0, 43)
    assert sum(len(p["gid"]) for p in parts) == 10000
    assert max(len(p["gid"]) for p in parts) < 1.3 * 2500
    # degenerate: everything in one row still yields valid, strictly increasing bounds
    b2 = slabs.partition_rows(np.full(100, 5.0, np.float32), 43, 60.0, 8)
    assert b2[0] == 0 and b2[-1] == 43 and all(b2[i] < b2[i + 1] for i in range(8))
    assert abs(slabs.halo_width(30.0) - (np.sqrt(np.float32(30)) * np.float32(1.05) + 0.25)) < 1e-5
