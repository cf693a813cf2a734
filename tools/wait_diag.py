"""Per-rank frame time of coupled slabs, of the same slabs without any agent near the boundary (RTS_DIAG_GAP=units), and of the
same half-maps as independent one-rank slabs (RTS_DIAG_SOLO=1, optionally RTS_DIAG_SOLO_COMM=1) — diagnostic. With a debug build
that exports rts_debug_wait_ns (k_ingest_remote clocking its spin into SlabDev::overflow[2..3]; not part of the product library)
it also prints the time actually spent waiting for the neighbour's word.
    torchrun --nproc-per-node 2 tools/wait_diag.py"""
import os, sys, json, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from rtsengine_b200 import engine, scenario, slabs, abi
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tile = scenario.load_map("map_c3.npz")
fx = scenario.tile_map(tile, 1, world)
nx, ny = (int(v) for v in fx["n_cells"])
base = engine.default_params(nx, ny); base.precision = 1
agents = scenario.tiled_agents(tile, 1, world, 1_000_000, seed=1236)
if os.environ.get("RTS_DIAG_GAP"):   # no agent within GAP units of a slab boundary: no halo traffic, no ghosts, no migrants
    gap = float(os.environ["RTS_DIAG_GAP"])
    yb = (agents["r"][:, 1] + 5120.0 / 2) % 5120.0 - 5120.0 / 2      # distance to the nearest multiple of the tile height
    keep = (np.abs(yb) > gap) | (agents["r"][:, 1] < 2560.0) | (agents["r"][:, 1] > 5120.0 * world - 2560.0)
    agents = {k: v[keep] for k, v in agents.items()}
agents["gid"] = np.arange(len(agents["r"]), dtype=np.uint32)
bounds = slabs.partition_rows(agents["r"][:, 1], base.ns_ny, base.ns_cell_size, world)
parts = slabs.split_agents(agents, bounds, base.ns_cell_size, base.ns_ny)
solo = os.environ.get("RTS_DIAG_SOLO") == "1"   # every rank a one-rank slab over its own rows: no neighbour, frames from the graph
if solo:
    base.reserved[5] = 16
    sr = slabs.SlabRank(base, [bounds[rank], bounds[rank + 1]], 0, local, cap_migrants=2344, cap_halo=7032)
    if os.environ.get("RTS_DIAG_SOLO_COMM") == "1":   # ... with an (idle) NCCL communicator of its own inside the library
        sr.init_comm(slabs.make_unique_id(sr.g.L))
else:
    sr = slabs.SlabRank(base, bounds, rank, local, cap_migrants=2344, cap_halo=7032)
    uid = [slabs.make_unique_id(sr.g.L) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0); sr.init_comm(uid[0])
sr.g.upload_map(scenario.tables_of(fx)); sr.g.upload_paths(scenario.paths_of(fx)); sr.g.set_agents(parts[rank])
if not solo: sr.halo_refresh()
sr.g.update(16, sync=True)
L = sr.g.L
has_dbg = hasattr(L, "rts_debug_wait_ns")
if has_dbg:
    L.rts_debug_wait_ns.restype = C.c_uint32; L.rts_debug_wait_ns.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for d in (0, 1): L.rts_debug_wait_ns(sr.g.h, d, 1)
dist.barrier(); torch.cuda.synchronize()
sr.g.update(8); sr.g.reset_timings(); F = 80; sr.g.update(F); sr.g.sync()
ms = sr.g.timings().last_step_ms / F
w = [L.rts_debug_wait_ns(sr.g.h, d, 0) / (F + 8) / 1000.0 for d in (0, 1)] if has_dbg else [float('nan')] * 2
for r in range(world):
    dist.barrier()
    if r == rank: print(json.dumps({"solo": solo, "gap": os.environ.get("RTS_DIAG_GAP"), "owned": len(parts[rank]["gid"]), "rank": rank, "ms_per_frame": round(ms, 5), "wait_us_per_frame_dir0_dir1": [round(x, 2) for x in w]}), flush=True)
sr.close(); dist.destroy_process_group()
