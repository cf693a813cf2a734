"""Per-rank timing of the slab path (diagnostic). torchrun --nproc-per-node N tools/slab_trace.py [agents_per_gpu] [frames] [cap_halo] [cap_mig]
RTS_TRACE_CONFIG4=1: BASELINE config 4 instead (8x8 config-3 tiles, agents_per_gpu * N agents in total, strong-scaling layout);
RTS_TRACE_PRECISION / RTS_RESERVED5 select the arithmetic mode and the exchange switches."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from rtsengine_b200 import engine, scenario, slabs

per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cap_halo = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 17
cap_mig = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 15
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tile = scenario.load_map("map_c3.npz")
c4 = os.environ.get("RTS_TRACE_CONFIG4") == "1"
fx = scenario.tile_map(tile, 8, 8) if c4 else scenario.tile_map(tile, 1, world)
nx, ny = (int(v) for v in fx["n_cells"])
base = engine.default_params(nx, ny)
base.precision = int(os.environ.get("RTS_TRACE_PRECISION", "1"))
base.reserved[5] = int(os.environ.get("RTS_RESERVED5", "0"))
agents = scenario.tiled_agents(tile, 8, 8, per_gpu * world // 64, seed=1237) if c4 else scenario.tiled_agents(tile, 1, world, per_gpu, seed=1236)
agents["gid"] = np.arange(len(agents["r"]), dtype=np.uint32)
if c4 and len(sys.argv) <= 3:   # capacities as the bench leg sizes them (slabs._slab_case)
    per_unit_height = len(agents["r"]) / (ny * 5.0)
    cap_halo = int(max(4096, 6 * per_unit_height * slabs.halo_width(base.r_max2)))
    cap_mig = int(max(1024, 6 * per_unit_height * 2.0))
bounds = slabs.partition_rows(agents["r"][:, 1], base.ns_ny, base.ns_cell_size, world)
parts = slabs.split_agents(agents, bounds, base.ns_cell_size, base.ns_ny)
sr = slabs.SlabRank(base, bounds, rank, local, cap_migrants=cap_mig, cap_halo=cap_halo)
uid = [slabs.make_unique_id(sr.g.L) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
sr.init_comm(uid[0])
sr.g.upload_map(scenario.tables_of(fx)); sr.g.upload_paths(scenario.paths_of(fx)); sr.g.set_agents(parts[rank]); sr.halo_refresh()
sr.g.update(10, sync=True)
res = {}
for mode, nf in (("profiled", frames), ("plain20_a", 20), ("plain20_b", 20), ("plain20_c", 20), ("plain", frames), ("plain_b", frames), ("profiled2", frames)):
    sr.g.set_profiling(mode.startswith("profiled"))
    dist.barrier(); torch.cuda.synchronize()
    t_enter = time.perf_counter()
    sr.g.reset_timings()
    t0 = time.perf_counter()
    sr.g.update(nf); t1 = time.perf_counter()
    sr.g.sync(); t2 = time.perf_counter()
    tt = torch.tensor([t_enter], dtype=torch.float64, device="cuda")   # (CLOCK_MONOTONIC is shared by the ranks of one box)
    lo, hi = tt.clone(), tt.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res[mode] = {"frames": nf, "enqueue_ms_per_frame": round(1e3 * (t1 - t0) / nf, 5), "wall_ms_per_frame": round(1e3 * (t2 - t0) / nf, 5),
                 "dev_ms_per_frame": round(sr.g.timings().last_step_ms / nf, 5), "start_skew_ms_over_ranks": round(1e3 * float((hi - lo).item()), 4)}
    if mode.startswith("profiled"):
        res[mode]["kernels"] = {k: round(sr.g.kernel_time_ms(k)[0], 4) for k in ("step", "dense", "scan", "reorder", "exchange", "ingest")}
own = sr.owned(("r",))
res["owned"] = len(own["gid"]); res["rank"] = rank; res["config4"] = c4; res["caps"] = [cap_halo, cap_mig]; res["fine_cells"] = [int(v) for v in sr.g.fine_grid()] if hasattr(sr.g, "fine_grid") else None
for r in range(world):
    dist.barrier()
    if r == rank: print(json.dumps(res), flush=True)
sr.close(); dist.destroy_process_group()
