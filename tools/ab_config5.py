"""config 5 on one GPU (diagnostic): whole-frame time and per-kernel times, 100 frames.  python tools/ab_config5.py [n]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario, abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
fx = scenario.load_map("map_c5.npz")
agents, paths = scenario.config5_agents(fx, n=n, seed=1238)
p = engine.default_params(2048, 2048)
p.precision = abi.PRECISION_FAST
with engine.GpuAgentSystem(p) as g:
    g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
    g.update(8, sync=True)
    g.reset_timings(); g.update(96); g.sync()
    ms = g.timings().last_step_ms / 96
    g.set_agents(agents); g.update(8, sync=True)
    g.set_profiling(True); g.reset_timings(); g.update(96, sync=True)
    kt = {k: round(g.kernel_time_ms(k)[0], 4) for k in ("step", "dense", "scan", "reorder", "walk")}
print(json.dumps({"lib": os.path.basename(engine.LIB_PATH), "agents": n, "ms_per_frame": round(ms, 4), "kernel_ms": kt}), flush=True)
