"""CPU: the C-ABI shared library loads, exports every symbol include/rts_gpu.h declares, its struct
layouts match the ctypes mirror, and it refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from rtsengine_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rts_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cuda_lib, n), f"{n} is declared in include/rts_gpu.h but not exported by librts_b200.so"
    # and the binding covers all of them
    assert set(names) <= set(cuda_lib._rts_signatures), set(names) - set(cuda_lib._rts_signatures)


def test_struct_layouts_match(cuda_lib):
    for which, st in enumerate([abi.RtsParams, abi.RtsTriangle, abi.RtsEdge, abi.RtsWallTable, abi.RtsPathTable,
                                abi.RtsUnitType, abi.RtsAgentView, abi.RtsTimings]):
        assert cuda_lib.rts_abi_sizeof(which) == C.sizeof(st), st.__name__
    assert abi.TRI_DTYPE.itemsize == C.sizeof(abi.RtsTriangle)
    assert abi.EDGE_DTYPE.itemsize == C.sizeof(abi.RtsEdge)
    assert abi.UNIT_DTYPE.itemsize == C.sizeof(abi.RtsUnitType)


def test_default_params_follow_the_reference(cuda_lib):
    p = engine.default_params(512, 512)
    # NeighbourSearcherContext.cpp:70-71 with box 2560, cell 60 -> 43; main.cpp:50 -> 64 cells of 40
    assert (p.ns_nx, p.ns_ny, p.ns_cell_size, p.r_max2) == (43, 43, 60.0, 30.0)
    assert (p.tri_nx, p.tri_ny, p.tri_cell_size) == (64, 64, 40.0)
    # UI.cpp:124-131
    assert [round(x, 6) for x in (p.mul_push, p.mul_repulse, p.mul_scatter, p.mul_decay, p.mul_velocity)] == [0.1, 1.0, 0.1, 1.0, 2.0]
    p2 = engine.default_params(1024, 1024)
    assert (p2.ns_nx, p2.tri_nx) == (86, 128)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback(cuda_lib):
    with pytest.raises(engine.RtsError) as e:
        engine.GpuAgentSystem(map_cells=(512, 512))
    assert e.value.code == abi.RTS_E_CUDA


def test_apply_delta_writes_the_change_list_into_the_blackboard(cuda_lib):
    """rts_apply_delta is host-only: entries {index, angle_vel, target, state} land in component order, absent arrays are
    skipped, an index beyond the population is refused."""
    import numpy as np
    n = 50
    assert cuda_lib.rts_abi_sizeof(8) == abi.DELTA_DTYPE.itemsize == 20
    ch = np.zeros(7, abi.DELTA_DTYPE)
    ch["index"] = [3, 49, 0, 17, 3, 20, 21]            # (3 twice: the later entry wins)
    ch["angle_vel"] = [1.5, -2.0, 0.25, 5.0, 7.0, 0.0, -0.5]
    ch["target_x"] = np.arange(7) * 10.0
    ch["target_y"] = np.arange(7) * -3.0
    ch["target_x"][5] = np.nan
    ch["state"] = [1, 0, 2, 1, 0, 1, 2]
    av, st, tg = np.full(n, 9.0, np.float32), np.full(n, 7, np.uint8), np.full((n, 2), 8.0, np.float32)
    assert cuda_lib.rts_apply_delta(ch.ctypes.data, 7, n, av.ctypes.data, st.ctypes.data, tg.ctypes.data) == abi.RTS_OK
    want_av, want_st, want_tg = np.full(n, 9.0, np.float32), np.full(n, 7, np.uint8), np.full((n, 2), 8.0, np.float32)
    for c in ch:
        want_av[c["index"]] = c["angle_vel"]; want_st[c["index"]] = c["state"]; want_tg[c["index"]] = (c["target_x"], c["target_y"])
    assert np.array_equal(av, want_av) and np.array_equal(st, want_st)
    assert np.array_equal(tg.view(np.uint32), want_tg.view(np.uint32))
    av2 = np.zeros(n, np.float32)
    assert cuda_lib.rts_apply_delta(ch.ctypes.data, 7, n, av2.ctypes.data, None, None) == abi.RTS_OK and av2[17] == 5.0
    ch["index"][6] = n
    assert cuda_lib.rts_apply_delta(ch.ctypes.data, 7, n, av.ctypes.data, st.ctypes.data, tg.ctypes.data) == abi.RTS_E_ARG


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """the drop-in boundary is a C ABI: include/rts_gpu.h compiles as C99 (-pedantic, no C++, no torch types) and a plain C
    program links against librts_b200.so and gets the no-device error through the header's own prototypes"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "c_abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "rts_gpu.h"\n'
        "int main(void) {\n"
        "    RtsParams p; RtsHandle h = 0;\n"
        "    rts_default_params(&p, 512, 512);\n"
        "    int32_t rc = rts_create(&p, 0, &h);\n"
        '    printf("%d %d %d\\n", (int)rc, (int)p.ns_nx, (int)sizeof(RtsAgentView));\n'
        "    if (rc == RTS_OK) rts_destroy(h);\n"
        "    return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "rtsengine_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)])
    exe = tmp_path / "c_abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", inc, str(src), "-L", libdir, "-lrts_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rc, ns_nx, size = (int(v) for v in out.stdout.split())
    assert ns_nx == 43 and size == C.sizeof(abi.RtsAgentView)
    if not torch.cuda.is_available():
        assert rc == abi.RTS_E_CUDA      # no device: an error code, never a CPU path
