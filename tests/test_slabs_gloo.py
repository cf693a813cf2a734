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


def _worker(rank, world, port_no, frames, result_dir):
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
    bounds = slabs.partition_rows(a["r"][:, 1], P.ns_ny, P.ns_cell_size, world)
    parts = slabs.split_agents(a, bounds, P.ns_cell_size, P.ns_ny)
    m = slabs.SlabModel(P, bounds[rank], bounds[rank + 1], lambda ag, gh: ow.step(ag, ghost=gh))
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
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), **m.owned)
    dist.destroy_process_group()


def test_slab_protocol_world_size_2(tmp_path):
    frames = 10
    port_no = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port_no, frames, str(tmp_path)), nprocs=2, join=True)
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
    parts = [dict(np.load(os.path.join(tmp_path, f"rank{r}.npz"))) for r in range(2)]
    gid = np.concatenate([p["gid"] for p in parts])
    assert len(gid) == 6000 and len(np.unique(gid)) == 6000
    order = np.argsort(gid)
    for k in ("r", "inertia", "angle", "angle_vel", "state", "waypoint", "r_next"):
        merged = np.concatenate([p[k] for p in parts])[order]
        H.assert_same_bits(merged, a[k], k)
    # the cut was actually crossed
    from rtsengine_b200 import slabs
    P = port.default_params(512, 512)
    bounds = slabs.partition_rows(r0[:, 1], P.ns_ny, P.ns_cell_size, 2)
    start_rows = slabs.row_of(r0[:, 1], P.ns_cell_size, P.ns_ny)
    end_rows = slabs.row_of(a["r"][:, 1], P.ns_cell_size, P.ns_ny)
    assert ((start_rows < bounds[1]) != (end_rows < bounds[1])).sum() > 0


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
