"""Exact vs tolerance mode on config 3 (diagnostic): whole-frame device time over three frame windows and the per-kernel
times (profiling mode, kernels serialised) over the same windows.   python tools/ab_modes.py [n] [modes e.g. 01]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtsengine_b200 import engine, scenario
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
modes = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "01")]
fx = scenario.load_map("map_c3.npz")
agents, paths = scenario.config3_agents(fx, n=n)
tables = scenario.tables_of(fx)
WINDOWS = [(5, 20), (25, 95), (120, 100)]          # (frames to skip from the previous window's end..., frames timed)
for prec in modes:
    p = engine.default_params(1024, 1024)
    p.precision = prec
    if os.environ.get("RTS_FINE_CS"):   # explicit gather-cell size (RtsParams.reserved[1] = the float's bits)
        import struct
        p.reserved[1] = struct.unpack("I", struct.pack("f", float(os.environ["RTS_FINE_CS"])))[0]
    p.reserved[5] = int(os.environ.get("RTS_RESERVED5", "0"))   # e.g. 64 = no programmatic dependent launch
    res = {"lib": os.path.basename(engine.LIB_PATH), "precision": prec, "reserved5": int(p.reserved[5]), "fine_cs": os.environ.get("RTS_FINE_CS", "auto")}
    with engine.GpuAgentSystem(p) as g:
        g.upload_map(tables); g.upload_paths(paths)
        for profiling in (False, True):
            g.set_agents(agents)
            g.set_profiling(profiling)
            frame = 0
            for start, count in WINDOWS:
                g.update(start - frame, sync=True)
                g.reset_timings(); g.update(count); g.sync()
                frame = start + count
                key = f"f{start}-{frame}"
                if not profiling:
                    res[key] = round(g.timings().last_step_ms / count, 5)
                else:
                    res[key + "_k"] = {k: round(g.kernel_time_ms(k)[0], 5) for k in ("step", "dense", "scan", "reorder", "walk")}
            g.set_profiling(False)
    print(json.dumps(res), flush=True)
