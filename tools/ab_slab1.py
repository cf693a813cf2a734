"""k_step of a slab handle against k_step of a single-domain handle on ONE GPU (diagnostic): a one-rank "decomposition" (no
neighbours, nothing is packed) runs the MULTI instantiation of the frame kernels over the same agents.
    python tools/ab_slab1.py [c3|c4] [n]      c4: the lowest quarter of BASELINE config 4 (what rank 0 of 4 owns)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtsengine_b200 import engine, scenario, slabs, abi
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
tile = scenario.load_map("map_c3.npz")
if which == "c3":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    fx = tile
    agents, paths = scenario.config3_agents(fx, n=n)
    cells = (1024, 1024)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    fx = scenario.tile_map(tile, 8, 8)
    agents = scenario.tiled_agents(tile, 8, 8, n // 64, seed=1237)
    paths = scenario.paths_of(fx)
    cells = (8192, 8192)
    parts = int(os.environ.get("RTS_SLAB_PARTS", "4"))         # rank 0 of `parts`
    keep = agents["r"][:, 1] < (683 // parts) * 60.0 - 100.0   # ... away from the boundary
    agents = {k: v[keep] for k, v in agents.items()}
agents["gid"] = np.arange(len(agents["r"]), dtype=np.uint32)
tables = scenario.tables_of(fx)
base = engine.default_params(*cells)
base.precision = abi.PRECISION_FAST
base.reserved[5] = int(os.environ.get("RTS_RESERVED5", "0"))   # e.g. 16: the one-rank slab's frames from a CUDA graph too
if os.environ.get("RTS_FINE_CS"):   # explicit gather-cell size
    import struct
    base.reserved[1] = struct.unpack("I", struct.pack("f", float(os.environ["RTS_FINE_CS"])))[0]
F0, F1 = 10, 60
def run(g, label):
    g.upload_map(tables); g.upload_paths(paths)
    res = {"handle": label, "agents": len(agents["r"]), "fine_cs": os.environ.get("RTS_FINE_CS", "auto")}
    for profiling in (False, True):
        g.set_agents(agents if label == "slab" else {k: v for k, v in agents.items() if k != "gid"})
        g.set_profiling(profiling)
        g.update(F0, sync=True)
        g.reset_timings(); g.update(F1); g.sync()
        if not profiling: res["ms_per_frame"] = round(g.timings().last_step_ms / F1, 5)
        else: res["kernel_ms"] = {k: round(g.kernel_time_ms(k)[0], 5) for k in ("step", "dense", "scan", "reorder", "ingest", "walk")}
    g.set_profiling(False)
    print(json.dumps(res), flush=True)
with engine.GpuAgentSystem(base) as g:
    run(g, "single")
rows_hi = int(base.ns_ny) if which == "c3" else 683 // int(os.environ.get("RTS_SLAB_PARTS", "4"))   # c4: the slab's own gather grid
sr = slabs.SlabRank(base, [0, rows_hi], 0, 0, cap_migrants=4096, cap_halo=16384)
run(sr.g, "slab")
sr.close()
