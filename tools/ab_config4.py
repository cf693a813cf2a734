"""config 4 on one GPU (diagnostic): whole-frame time and per-kernel times for several fine-cell sizes.  python tools/ab_config4.py [n] [cs ...]"""
import sys, os, json, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario, abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sizes = [float(x) for x in sys.argv[2:]] or [6.0, 9.0, 12.0]
tile = scenario.load_map("map_c3.npz")
fx = scenario.tile_map(tile, 8, 8)
agents = scenario.tiled_agents(tile, 8, 8, n // 64, seed=1237)
tables, paths = scenario.tables_of(fx), scenario.paths_of(fx)
for cs in sizes:
    p = engine.default_params(8192, 8192)
    p.precision = abi.PRECISION_FAST
    p.reserved[1] = struct.unpack("I", struct.pack("f", cs))[0] if cs > 0 else 0
    with engine.GpuAgentSystem(p) as g:
        g.upload_map(tables); g.upload_paths(paths); g.set_agents(agents)
        g.update(5, sync=True)
        g.reset_timings(); g.update(40); g.sync()
        ms = g.timings().last_step_ms / 40
        g.set_agents(agents); g.update(5, sync=True)
        g.set_profiling(True); g.reset_timings(); g.update(40, sync=True)
        kt = {k: round(g.kernel_time_ms(k)[0], 4) for k in ("step", "dense", "scan", "reorder", "walk")}
    print(json.dumps({"fine_cs": cs, "agents": len(agents["r"]), "ms_per_frame": round(ms, 4), "kernel_ms": kt}), flush=True)
