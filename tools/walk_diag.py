"""Which agents make the triangle walk slow on config 5 (diagnostic): positions after `frames` frames, where they lie, and
the time of the stand-alone query (rts_find_triangles = the reference walk from the cache cell, no hint) on subsets.
    python tools/walk_diag.py [n] [frames]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rtsengine_b200 import engine, scenario, abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
fx = scenario.load_map("map_c5.npz")
agents, paths = scenario.config5_agents(fx, n=n, seed=1238)
nx, ny = (int(v) for v in fx["n_cells"])
W, H = nx * 5.0, ny * 5.0
p = engine.default_params(nx, ny)
p.precision = abi.PRECISION_FAST
with engine.GpuAgentSystem(p) as g:
    g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
    g.update(frames, sync=True)
    out = g.communicate(("r", "tri_idx", "state"))
    r = out["r"]
    fin = np.isfinite(r).all(1)
    outside = fin & ((r[:, 0] < 0) | (r[:, 0] >= W) | (r[:, 1] < 0) | (r[:, 1] >= H))
    far = fin & ((r[:, 0] < -40) | (r[:, 0] >= W + 40) | (r[:, 1] < -40) | (r[:, 1] >= H + 40))
    res = {"agents": n, "frames": frames, "not_finite": int((~fin).sum()), "outside_map": int(outside.sum()), "outside_by_more_than_a_cache_cell": int(far.sum()),
           "x_range": [float(r[fin, 0].min()), float(r[fin, 0].max())], "y_range": [float(r[fin, 1].min()), float(r[fin, 1].max())]}
    def timed(sel, label):
        q = np.ascontiguousarray(r[sel][:20000])
        if len(q) == 0:
            return
        g.find_triangles(q[:64])
        t0 = time.perf_counter(); t = g.find_triangles(q); dt = time.perf_counter() - t0
        res[label] = {"queries": len(q), "ms": round(1e3 * dt, 3), "distinct_triangles": int(len(np.unique(t)))}
        for k in (1, 32, 1024):   # how the time depends on the number of queries: latency of the longest walk vs throughput
            if len(q) >= k:
                t0 = time.perf_counter(); g.find_triangles(q[:k]); res[label][f"ms_first_{k}"] = round(1e3 * (time.perf_counter() - t0), 3)
    timed(fin & ~outside, "inside")
    timed(outside & ~far, "just_outside")
    timed(far, "far_outside")
    # the slowest single queries among the outside ones
    idx = np.nonzero(outside)[0][:200]
    slow = []
    for i in idx[:60]:
        t0 = time.perf_counter(); g.find_triangles(r[i:i + 1]); slow.append((time.perf_counter() - t0, i))
    slow.sort(reverse=True)
    res["slowest_single_queries_ms"] = [[round(1e3 * dt, 3), r[i].tolist()] for dt, i in slow[:5]]
print(json.dumps(res), flush=True)
