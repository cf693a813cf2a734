"""Frame and kernel times of BASELINE config 4 restricted to the lowest 1/k of the map (what one of k slabs owns), single
handle, one GPU (diagnostic: how the per-agent cost changes with the population).  python tools/ab_c4_sizes.py [k ...]
RTS_C4_ONLY=k: just that population, 30 frames, no profiling (the target of an ncu capture)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtsengine_b200 import engine, scenario, abi
ks = [int(x) for x in sys.argv[1:]] or [8, 4, 2, 1]
only = int(os.environ.get("RTS_C4_ONLY", "0"))
tile = scenario.load_map("map_c3.npz")
fx = scenario.tile_map(tile, 8, 8)
full = scenario.tiled_agents(tile, 8, 8, 10_000_000 // 64, seed=1237)
tables, paths = scenario.tables_of(fx), scenario.paths_of(fx)
H = 8192 * 5.0
for k in ([only] if only else ks):
    keep = full["r"][:, 1] < H / k
    agents = {n: v[keep] for n, v in full.items()}
    p = engine.default_params(8192, 8192)
    p.precision = abi.PRECISION_FAST
    if only: p.reserved[2] = 1
    with engine.GpuAgentSystem(p) as g:
        g.upload_map(tables); g.upload_paths(paths); g.set_agents(agents)
        if only:
            g.update(30, sync=True); print("done", g.n); continue
        g.update(10, sync=True)
        g.reset_timings(); g.update(40); g.sync()
        ms = g.timings().last_step_ms / 40
        g.set_agents(agents); g.update(10, sync=True)
        g.set_profiling(True); g.reset_timings(); g.update(40, sync=True)
        kt = {n: round(g.kernel_time_ms(n)[0], 5) for n in ("step", "dense", "scan", "reorder")}
    print(json.dumps({"slabs": k, "agents": int(keep.sum()), "ms_per_frame": round(ms, 5), "kernel_ms": kt,
                      "step_ns_per_agent": round(1e6 * kt["step"] / keep.sum(), 2)}), flush=True)
