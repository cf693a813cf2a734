"""ctypes mirror of include/rts_gpu.h (plain declarations, no behaviour).

The struct layouts here must match the C header field for field; tests/test_abi.py checks the
sizes against `rts_abi_sizeof` exported by the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

RTS_OK = 0
RTS_E_CUDA, RTS_E_NCCL, RTS_E_ARG, RTS_E_CAPACITY, RTS_E_STATE = -1, -2, -3, -4, -5
MOVING, STANDING, HOLDING = 0, 1, 2
FLAG_REPLAN, FLAG_WALK_FAILED = 1, 2
PRECISION_EXACT, PRECISION_FAST = 0, 1
NO_TRI = 0xFFFFFFFF

STATUS_NAMES = {0: "RTS_OK", -1: "RTS_E_CUDA", -2: "RTS_E_NCCL", -3: "RTS_E_ARG", -4: "RTS_E_CAPACITY",
                -5: "RTS_E_STATE"}


class RtsParams(C.Structure):
    _fields_ = [
        ("map_nx", C.c_int32), ("map_ny", C.c_int32), ("map_cell_size", C.c_float),
        ("ns_nx", C.c_int32), ("ns_ny", C.c_int32), ("ns_cell_size", C.c_float), ("r_max2", C.c_float),
        ("tri_nx", C.c_int32), ("tri_ny", C.c_int32), ("tri_cell_size", C.c_float),
        ("mul_push", C.c_float), ("mul_repulse", C.c_float), ("mul_scatter", C.c_float), ("mul_align", C.c_float),
        ("mul_seek", C.c_float), ("mul_decay", C.c_float), ("mul_velocity", C.c_float),
        ("dt", C.c_float), ("rhard", C.c_float), ("rscatter", C.c_float), ("sys_max_speed", C.c_float),
        ("finish_angle", C.c_float), ("finish_radius", C.c_float), ("finish_n_steps", C.c_float),
        ("finish_cone_len", C.c_float),
        ("enable_arrival", C.c_uint32), ("walk_all", C.c_uint32),
        ("slab_row_lo", C.c_int32), ("slab_row_hi", C.c_int32),
        ("reserved", C.c_uint32 * 6),
        ("precision", C.c_uint32),
        ("enable_alignment", C.c_uint32), ("ralign", C.c_float),
    ]


class RtsTriangle(C.Structure):
    _fields_ = [("vx", C.c_int32 * 3), ("vy", C.c_int32 * 3), ("neighbours", C.c_uint32 * 3),
                ("is_constrained", C.c_uint8 * 3), ("type", C.c_uint8), ("pad_", C.c_uint32 * 2)]


TRI_DTYPE = np.dtype([("vx", np.int32, 3), ("vy", np.int32, 3), ("neighbours", np.uint32, 3),
                      ("is_constrained", np.uint8, 3), ("type", np.uint8), ("pad_", np.uint32, 2)])
EDGE_DTYPE = np.dtype([("from_x", np.float32), ("from_y", np.float32), ("t_x", np.float32), ("t_y", np.float32),
                       ("l", np.float32)])
UNIT_DTYPE = np.dtype([("lambda_", np.float32), ("max_speed", np.float32), ("seek_max_speed", np.float32),
                       ("turn_rate", np.float32), ("desired_angle", np.float32), ("pad_", np.float32, 3)])
assert TRI_DTYPE.itemsize == 48 and EDGE_DTYPE.itemsize == 20 and UNIT_DTYPE.itemsize == 32


class RtsEdge(C.Structure):
    _fields_ = [("from_x", C.c_float), ("from_y", C.c_float), ("t_x", C.c_float), ("t_y", C.c_float), ("l", C.c_float)]


class RtsWallTable(C.Structure):
    _fields_ = [("edges", C.c_void_p * 4), ("line_offsets", C.c_void_p * 4), ("n_lines", C.c_uint32 * 4)]


class RtsPathTable(C.Structure):
    _fields_ = [("n_paths", C.c_uint32), ("offsets", C.c_void_p), ("waypoints", C.c_void_p), ("portals", C.c_void_p),
                ("path_end", C.c_void_p), ("group", C.c_void_p), ("n_groups", C.c_uint32)]


class RtsUnitType(C.Structure):
    _fields_ = [("lambda_", C.c_float), ("max_speed", C.c_float), ("seek_max_speed", C.c_float),
                ("turn_rate", C.c_float), ("desired_angle", C.c_float), ("pad_", C.c_float * 3)]


AGENT_FIELDS = [  # name, numpy dtype, components per agent
    ("r", np.float32, 2), ("vel", np.float32, 2), ("inertia", np.float32, 2), ("angle", np.float32, 1),
    ("angle_vel", np.float32, 1), ("radius", np.float32, 1), ("mass", np.float32, 1), ("state", np.uint8, 1),
    ("player", np.uint8, 1), ("unit_type", np.uint8, 1), ("path_id", np.int32, 1), ("waypoint", np.int32, 1),
    ("r_next", np.float32, 2), ("gid", np.uint32, 1), ("tri_idx", np.uint32, 1), ("ns_cell", np.uint32, 1),
    ("map_cell", np.uint32, 1), ("flags", np.uint32, 1),
]
DOWNLOAD_ONLY = {"vel", "tri_idx", "ns_cell", "map_cell", "flags"}


class RtsAgentView(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in AGENT_FIELDS]


class RtsTimings(C.Structure):
    _fields_ = [("last_step_ms", C.c_double), ("kernel_ms", C.c_double * 8), ("kernel_launches", C.c_uint64),
                ("steps", C.c_uint64)]


class RtsExchangeBuf(C.Structure):
    _fields_ = [("data_dev", C.c_void_p), ("count_dev", C.c_void_p), ("capacity", C.c_uint32),
                ("record_bytes", C.c_uint32)]


def agent_view(arrays: dict, n: int, keep: list) -> RtsAgentView:
    """RtsAgentView over numpy arrays (checked for dtype / shape / contiguity). `keep` receives the
    arrays so they outlive the call."""
    v = RtsAgentView()
    for name, dt, k in AGENT_FIELDS:
        a = arrays.get(name)
        if a is None:
            setattr(v, name, None)
            continue
        a = np.ascontiguousarray(a, dtype=dt)
        if a.size != n * k:
            raise ValueError(f"agent array {name!r}: expected {n * k} elements, got {a.size}")
        keep.append(a)
        setattr(v, name, a.ctypes.data)
    return v


def alloc_agent_arrays(n: int, names) -> dict:
    out = {}
    for name, dt, k in AGENT_FIELDS:
        if name in names:
            out[name] = np.zeros((n, k) if k > 1 else (n,), dt)
    return out


def wall_table(walls: dict, keep: list) -> RtsWallTable:
    """walls: {"edges0..3": (m,5) float32, "line_off0..3": uint32}"""
    w = RtsWallTable()
    for o in range(4):
        e = np.ascontiguousarray(walls[f"edges{o}"], np.float32).reshape(-1, 5)
        off = np.ascontiguousarray(walls[f"line_off{o}"], np.uint32)
        if e.shape[0] == 0:
            e = np.zeros((1, 5), np.float32)
        keep += [e, off]
        w.edges[o] = e.ctypes.data
        w.line_offsets[o] = off.ctypes.data
        w.n_lines[o] = len(off) - 1
    return w


def triangles_array(tri_verts, tri_nbrs, tri_constrained=None) -> np.ndarray:
    """(n,6) int32 [x0 y0 x1 y1 x2 y2] + (n,3) uint32 → RtsTriangle records"""
    n = len(tri_verts)
    t = np.zeros(n, TRI_DTYPE)
    tv = np.asarray(tri_verts, np.int32).reshape(n, 3, 2)
    t["vx"] = tv[:, :, 0]
    t["vy"] = tv[:, :, 1]
    t["neighbours"] = np.asarray(tri_nbrs, np.uint32)
    if tri_constrained is not None:
        t["is_constrained"] = np.asarray(tri_constrained, np.uint8)
    return t


def path_table(paths: dict, keep: list) -> RtsPathTable:
    """paths: {"offsets": uint32 (p+1), "waypoints": (w,2) f32, "portals": (w,5) f32, "path_end": (p,2) f32}"""
    off = np.ascontiguousarray(paths["offsets"], np.uint32)
    wp = np.ascontiguousarray(paths["waypoints"], np.float32).reshape(-1, 2)
    po = np.ascontiguousarray(paths["portals"], np.float32).reshape(-1, 5)
    pe = np.ascontiguousarray(paths["path_end"], np.float32).reshape(-1, 2)
    if len(wp) != off[-1] or len(po) != off[-1] or len(pe) != len(off) - 1:
        raise ValueError("path table arrays are inconsistent")
    if len(wp) == 0:
        wp = np.zeros((1, 2), np.float32)
        po = np.zeros((1, 5), np.float32)
    if len(pe) == 0:
        pe = np.zeros((1, 2), np.float32)
    keep += [off, wp, po, pe]
    t = RtsPathTable()
    t.n_paths = len(off) - 1
    t.offsets, t.waypoints, t.portals, t.path_end = off.ctypes.data, wp.ctypes.data, po.ctypes.data, pe.ctypes.data
    grp = paths.get("group")
    if grp is not None and len(grp):
        grp = np.ascontiguousarray(grp, np.int32)
        if len(grp) != t.n_paths:
            raise ValueError("path table: one group id per path")
        keep.append(grp)
        t.group = grp.ctypes.data
        t.n_groups = int(grp.max()) + 1
    else:
        t.group = None
        t.n_groups = t.n_paths
    return t


def default_unit_types() -> np.ndarray:
    """UnitTypes.txt: "default" (RADIUS 3, MASS 1) and "new unit" (RADIUS 5, MASS 50); the per-type
    dynamics constants are the component defaults (src/Systems/Components.h:84-119)."""
    u = np.zeros(2, UNIT_DTYPE)
    u["lambda_"] = 0.169
    u["max_speed"] = 25.0
    u["seek_max_speed"] = 50.0
    u["turn_rate"] = 5.0
    u["desired_angle"] = 0.0
    return u


# RtsAgentDelta (rts_download_delta), 20 bytes
DELTA_DTYPE = np.dtype([("index", np.uint32), ("angle_vel", np.float32), ("target_x", np.float32), ("target_y", np.float32), ("state", np.uint32)])
