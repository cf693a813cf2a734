"""Target of the ncu captures (diagnostic): config 3, `frames` frames launched kernel by kernel (no CUDA graph).
    python tools/ncu_target.py [precision 0|1] [frames] [n] [c5]
e.g.  ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 30 --launch-count 1 -o out python tools/ncu_target.py 1 40"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
c5 = len(sys.argv) > 4 and sys.argv[4] == "c5"      # BASELINE config 5 (corridor map) instead of config 3
fx = scenario.load_map("map_c5.npz" if c5 else "map_c3.npz")
agents, paths = scenario.config5_agents(fx, n=n, seed=1238) if c5 else scenario.config3_agents(fx, n=n)
p = engine.default_params(*(int(v) for v in fx["n_cells"]))
p.precision = prec
p.reserved[2] = 1          # no CUDA graph: every kernel is a launch ncu can pick
with engine.GpuAgentSystem(p) as g:
    g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
    g.update(frames, sync=True)
    print("done", g.n)
