"""k_step time over the first frames after an upload (diagnostic; used for timing experiments whose results are wrong after
a frame or two).  python tools/ab_first_frames.py [frames]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario, abi
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=1_000_000)
p = engine.default_params(1024, 1024)
p.precision = abi.PRECISION_FAST
with engine.GpuAgentSystem(p) as g:
    g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths)
    g.set_profiling(True)
    out = []
    for rep in range(4):
        g.set_agents(agents)
        g.reset_timings(); g.update(frames, sync=True)
        out.append({k: round(g.kernel_time_ms(k)[0], 5) for k in ("step", "dense", "scan", "reorder")})
print(json.dumps({"lib": os.path.basename(engine.LIB_PATH), "frames": frames, "kernel_ms": out}), flush=True)
