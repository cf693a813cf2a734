"""A/B of RtsParams.reserved[] switches on config 3 (diagnostic): whole-frame device time per variant, same frames."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=n)
tables = scenario.tables_of(fx)
only_default = len(sys.argv) > 2
for name, idx, val in ((("default", None, 0), ("default again", None, 0)) if only_default else (("default", None, 0), ("walk on main stream", 4, 1), ("no CUDA graph", 2, 1), ("default again", None, 0))):
    p = engine.default_params(1024, 1024)
    if idx is not None:
        p.reserved[idx] = val
    with engine.GpuAgentSystem(p) as g:
        g.upload_map(tables); g.upload_paths(paths); g.set_agents(agents)
        g.update(20, sync=True)
        out = []
        for w in range(2):
            g.reset_timings(); g.update(100); g.sync()
            out.append(round(g.timings().last_step_ms / 100, 5))
        print(json.dumps({"variant": name, "ms_per_frame": out}), flush=True)
