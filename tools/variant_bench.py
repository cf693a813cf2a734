"""A/B of k_step build variants (diagnostic): mean k_step ms over three 80-frame windows of config 3."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtsengine_b200 import engine, scenario
staged = int(sys.argv[1]) if len(sys.argv) > 1 else 0
windows = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=1_000_000)
p = engine.default_params(1024, 1024)
p.reserved[3] = staged
g = engine.GpuAgentSystem(p)
g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
g.set_profiling(True)
out = []
for s in range(windows):
    g.reset_timings(); g.update(80, sync=True)
    out.append(round(g.kernel_time_ms("step")[0], 4))
print(json.dumps({"staged": staged, "k_step_ms": out}))
