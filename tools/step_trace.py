"""k_step time and crowd density along a config-3 run (diagnostic, not a bench)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtsengine_b200 import engine, scenario

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
total = int(sys.argv[2]) if len(sys.argv) > 2 else 400
every = int(sys.argv[3]) if len(sys.argv) > 3 else 20
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=n)
g = engine.GpuAgentSystem(engine.default_params(1024, 1024))
g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
g.set_profiling(True)
for s in range(0, total, every):
    g.reset_timings()
    g.update(every, sync=True)
    ms = {k: g.kernel_time_ms(k)[0] for k in ("step", "scan", "reorder", "walk", "dense")}
    out = g.communicate(("r", "waypoint", "flags"))
    r = out["r"]
    # density: agents per 6x6 cell, max and 99.9 percentile
    c = (r[:, 0] // 6).astype(np.int64) * 4096 + (r[:, 1] // 6).astype(np.int64)
    cnt = np.bincount(c[(c >= 0)])
    cnt = cnt[cnt > 0]
    cc = (r[:, 0] // 60).astype(np.int64) * 4096 + (r[:, 1] // 60).astype(np.int64)
    ccnt = np.bincount(cc[cc >= 0])
    print(json.dumps({"frame": s + every, "max_per_ns_cell": int(ccnt.max()), "k_step_ms": round(ms["step"], 4), "scan_ms": round(ms["scan"], 4), "reorder_ms": round(ms["reorder"], 4), "walk_ms": round(ms["walk"], 4), "dense_ms": round(ms["dense"], 4),
                      "max_per_cell": int(cnt.max()), "p999": float(np.percentile(cnt, 99.9)), "mean_occ": float(cnt.mean()),
                      "nan": int(np.isnan(r).any(1).sum()), "wp_mean": float(out["waypoint"].mean())}), flush=True)
