"""Where the time of one end-to-end frame goes (diagnostic): update_shared_data_indexed / update / communicate (full) or
communicate_delta, host clock around each call, config 3.   python tools/e2e_breakdown.py [n]"""
import sys, os, json, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rtsengine_b200 import engine, scenario, abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
def pinned(shape, dtype):
    t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    return t, t.numpy()
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=n)
p = engine.default_params(1024, 1024)
p.precision = abi.PRECISION_FAST
keep = []
with engine.GpuAgentSystem(p) as g:
    g.upload_map(scenario.tables_of(fx)); g.upload_paths(paths); g.set_agents(agents)
    g.update(20, sync=True)
    board = {}
    for k, shp, dt in (("vel", (n, 2), np.float32), ("angle_vel", (n,), np.float32), ("state", (n,), np.uint8), ("r_next", (n, 2), np.float32)):
        t, a = pinned(shp, dt); keep.append(t); board[k] = a
    t, raw = pinned((n * abi.DELTA_DTYPE.itemsize,), np.uint8); keep.append(t)
    changes = raw.view(abi.DELTA_DTYPE)
    ptrs = {k: v.ctypes.data for k, v in board.items()}
    res = {}
    for mode in ("full", "delta", "delta_raw"):
        T = {"step": 0.0, "down": 0.0, "apply": 0.0}
        cnt = C.c_uint32(0)
        for f in range(33):
            t0 = time.perf_counter(); g.update(1); g.sync(); t1 = time.perf_counter()
            if mode == "full":
                g.communicate_ptrs(ptrs); t2 = t3 = time.perf_counter()
            elif mode == "delta":
                g.communicate_delta(board, changes); t2 = t3 = time.perf_counter()
            else:
                g._chk(g.L.rts_download_delta(g.h, ptrs["vel"], changes.ctypes.data, n, C.byref(cnt), None, None, None)); t2 = time.perf_counter()
                g.L.rts_apply_delta(changes.ctypes.data, cnt.value, n, ptrs["angle_vel"], ptrs["state"], ptrs["r_next"]); t3 = time.perf_counter()
            if f >= 3:
                T["step"] += t1 - t0; T["down"] += t2 - t1; T["apply"] += t3 - t2
        res[mode] = {k: round(1e3 * v / 30, 4) for k, v in T.items()}
        res[mode]["changed_last"] = int(cnt.value)
print(json.dumps(res), flush=True)
