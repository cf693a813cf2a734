"""TEST INFRASTRUCTURE — ctypes view of oracle/_ref/libref_harness*.so (verbatim reference TUs).

Only tests/, __graft_entry__.smoke(), tests/golden/make_golden.py and bench.py's CPU-baseline
legs may import this module.  The library is built by `make -C oracle ref` from the reference
sources where they lie under /root/reference (see oracle/Makefile); it travels to the GPU box as
a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def lib_path(fast: bool = False) -> str:
    return os.path.join(_HERE, "_ref", "libref_harness_fast.so" if fast else "libref_harness.so")


def available(fast: bool = False) -> bool:
    return os.path.exists(lib_path(fast))


_libs: dict = {}


def load(fast: bool = False):
    if fast in _libs:
        return _libs[fast]
    L = C.CDLL(lib_path(fast))
    vp = C.c_void_p
    L.refh_create.restype = vp
    L.refh_create.argtypes = [C.c_int, C.c_int]
    L.refh_destroy.argtypes = [vp]
    L.refh_last_error.restype = C.c_char_p
    L.refh_last_error.argtypes = [vp]
    L.refh_set_multiplier.argtypes = [vp, C.c_int, C.c_float]
    L.refh_add_building.argtypes = [vp] + [C.c_int] * 5
    L.refh_finalize_map.argtypes = [vp]
    L.refh_update_cell_grid.argtypes = [vp]
    L.refh_consistent.argtypes = [vp]
    L.refh_add_agent.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
    L.refh_n_agents.argtypes = [vp]
    L.refh_issue_paths.argtypes = [vp, i32p, C.c_int, C.c_float, C.c_float]
    L.refh_plan_now.argtypes = [vp]
    L.refh_step.argtypes = [vp, C.c_int]
    L.refh_get_state.argtypes = [vp] + [C.c_void_p] * 6
    L.refh_set_state.argtypes = [vp] + [C.c_void_p] * 5
    L.refh_set_phys_params.argtypes = [vp, C.c_int, C.c_float, C.c_float]
    L.refh_set_seek_params.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float]
    L.refh_get_seek_window.argtypes = [vp, f32p]
    L.refh_set_seek_window.argtypes = [vp, f32p]
    L.refh_group_of.argtypes = [vp, C.c_int]
    L.refh_n_triangles.argtypes = [vp]
    L.refh_get_triangles.argtypes = [vp, i32p, u32p, u8p]
    L.refh_tri_grid_cells.argtypes = [vp]
    L.refh_get_cell2tri.argtypes = [vp, u32p]
    L.refh_n_lines.argtypes = [vp, C.c_int]
    L.refh_n_edges.argtypes = [vp, C.c_int]
    L.refh_get_edges.argtypes = [vp, C.c_int, u32p, f32p]
    L.refh_find_triangles.argtypes = [vp, f32p, C.c_int, u32p]
    L.refh_find_triangle_int.restype = C.c_uint32
    L.refh_find_triangle_int.argtypes = [vp, C.c_int, C.c_int]
    L.refh_insert_vertex.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.refh_coord_to_cell.argtypes = [C.c_int, C.c_int, C.c_float, f32p, C.c_int, u64p]
    L.refh_physics_grid_nx.argtypes = [vp]
    L.refh_contact_edges.argtypes = [vp, C.c_float, C.c_float, C.c_float, f32p, C.c_int]
    L.refh_physics_neighbours.argtypes = [vp, C.c_int, i32p, C.c_int]
    L.refh_physics_ns_update.argtypes = [vp]
    L.refh_seek_ns_test.argtypes = [C.c_float, C.c_float, f32p, f32p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.refh_seek_ns_full.argtypes = [C.c_float, C.c_float, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_float]
    L.refh_plan_path.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p, C.c_int]
    _libs[fast] = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefWorld:
    """One headless reference world (Physics + Seek systems, CDT, Edges, planner)."""

    def __init__(self, n_cells_x: int = 512, n_cells_y: int = 512, fast: bool = False):
        self.L = load(fast)
        self.h = self.L.refh_create(n_cells_x, n_cells_y)
        self.n_cells = (n_cells_x, n_cells_y)

    def close(self):
        if self.h:
            self.L.refh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.refh_last_error(self.h).decode())
        return rc

    # map ------------------------------------------------------------------
    def add_building(self, cx, cy, n=8, m=8, c=1):
        return self.L.refh_add_building(self.h, cx, cy, n, m, c)

    def rebuild_cell_grid(self, last_found: int):
        """Triangulation::updateCellGrid with last_found_tri_ind_ preset (the cache then appears in tables())"""
        self.L.refh_rebuild_cell_grid.argtypes = [C.c_void_p, C.c_int]
        self.L.refh_rebuild_cell_grid.restype = None
        self.L.refh_rebuild_cell_grid(self.h, int(last_found))

    def last_found(self) -> int:
        self.L.refh_last_found.argtypes = [C.c_void_p]
        self.L.refh_last_found.restype = C.c_int
        return int(self.L.refh_last_found(self.h))

    def finalize_map(self):
        self._chk(self.L.refh_finalize_map(self.h))

    def update_cell_grid(self):
        self.L.refh_update_cell_grid(self.h)

    def consistent(self) -> bool:
        return bool(self.L.refh_consistent(self.h))

    def tables(self) -> dict:
        """CDT + wall tables in the flat layout the C ABI (include/rts_gpu.h) takes."""
        nt = self.L.refh_n_triangles(self.h)
        verts = np.zeros((nt, 6), np.int32)
        nbrs = np.zeros((nt, 3), np.uint32)
        cons = np.zeros((nt, 3), np.uint8)
        self.L.refh_get_triangles(self.h, verts, nbrs, cons)
        c2t = np.zeros(self.L.refh_tri_grid_cells(self.h), np.uint32)
        self.L.refh_get_cell2tri(self.h, c2t)
        out = {"tri_verts": verts, "tri_nbrs": nbrs, "tri_constrained": cons, "cell2tri": c2t,
               "n_cells": np.array(self.n_cells, np.int32)}
        for o in range(4):
            nl = self.L.refh_n_lines(self.h, o)
            ne = self.L.refh_n_edges(self.h, o)
            off = np.zeros(nl + 1, np.uint32)
            ed = np.zeros((max(ne, 1), 5), np.float32)
            self.L.refh_get_edges(self.h, o, off, ed)
            out[f"line_off{o}"] = off
            out[f"edges{o}"] = ed[:ne]
        return out

    # agents ---------------------------------------------------------------
    def add_agent(self, x, y, angle=0.0, radius=3.0, mass=1.0, player=0, state=1):
        return self._chk(self.L.refh_add_agent(self.h, x, y, angle, radius, mass, player, state))

    @property
    def n(self):
        return self.L.refh_n_agents(self.h)

    # PhysicsSystem::Multiplier (PhysicsSystem.h:21-29): the sliders of UI.cpp:121-131
    PUSH, REPULSE, SCATTER, ALIGN, SEEK, DECAY, VELOCITY = range(7)

    def set_multiplier(self, which: int, value: float):
        self.L.refh_set_multiplier(self.h, int(which), float(value))

    def set_unit_dynamics(self, i: int, lambda_: float, max_speed: float, seek_max_speed: float, turn_rate: float, desired_angle=0.0):
        """per-component constants (Components.h:84-119) of agent i"""
        self.L.refh_set_phys_params(self.h, int(i), float(lambda_), float(max_speed))
        self.L.refh_set_seek_params(self.h, int(i), float(seek_max_speed), float(turn_rate), float(desired_angle))

    def issue_paths(self, ids, tx, ty):
        ids = np.ascontiguousarray(ids, np.int32)
        self._chk(self.L.refh_issue_paths(self.h, ids, len(ids), tx, ty))

    def plan_now(self):
        self._chk(self.L.refh_plan_now(self.h))

    def step(self, n=1):
        self._chk(self.L.refh_step(self.h, n))

    def state(self) -> dict:
        n = self.n
        d = {"r": np.zeros((n, 2), np.float32), "vel": np.zeros((n, 2), np.float32),
             "inertia": np.zeros((n, 2), np.float32), "angle": np.zeros(n, np.float32),
             "angle_vel": np.zeros(n, np.float32), "state": np.zeros(n, np.int32)}
        self.L.refh_get_state(self.h, _ptr(d["r"]), _ptr(d["vel"]), _ptr(d["inertia"]), _ptr(d["angle"]),
                              _ptr(d["angle_vel"]), _ptr(d["state"]))
        return d

    def set_state(self, r=None, inertia=None, angle=None, angle_vel=None, state=None):
        cv = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        r, inertia, angle, angle_vel = (cv(a, np.float32) for a in (r, inertia, angle, angle_vel))
        state = cv(state, np.int32)
        self.L.refh_set_state(self.h, _ptr(r), _ptr(inertia), _ptr(angle), _ptr(angle_vel), _ptr(state))

    def seek_window(self):
        w = np.zeros((self.n, 16), np.float32)
        self.L.refh_get_seek_window(self.h, w)
        return w

    def set_seek_window(self, w):
        self.L.refh_set_seek_window(self.h, np.ascontiguousarray(w, np.float32))

    def groups(self):
        return np.array([self.L.refh_group_of(self.h, i) for i in range(self.n)], np.int32)

    # queries --------------------------------------------------------------
    def find_triangles(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.uint32)
        self.L.refh_find_triangles(self.h, xy, len(xy), out)
        return out

    def contact_edges(self, x, y, radius, cap=32):
        out = np.zeros((cap, 5), np.float32)
        k = self.L.refh_contact_edges(self.h, x, y, radius, out, cap)
        return out[:min(k, cap)]

    def physics_neighbours(self, i, cap=500):
        out = np.zeros(cap, np.int32)
        k = self.L.refh_physics_neighbours(self.h, i, out, cap)
        return out[:k]

    def plan_path(self, sx, sy, ex, ey, radius=3.0, cap=256):
        path = np.zeros((cap, 2), np.float32)
        portals = np.zeros((cap, 5), np.float32)
        k = self._chk(self.L.refh_plan_path(self.h, sx, sy, ex, ey, radius, path, portals, cap))
        return path[:k], portals[:k]


def attack_targets(box, cell, xy, player):
    """NeighbourSearcherStrategyAttack::execute through the reference's own context (AttackSystem.cpp:24): per agent the enemy
    index it reports, or -1"""
    L = load()
    L.refh_attack_ns.argtypes = [C.c_float, C.c_float, f32p, u8p, C.c_int, i32p]
    xy = np.ascontiguousarray(xy, np.float32)
    player = np.ascontiguousarray(player, np.uint8)
    out = np.zeros(len(xy), np.int32)
    L.refh_attack_ns(float(box), float(cell), xy, player, len(xy), out)
    return out


def coord_to_cell(nx, ny, cs, xy, fast=False):
    xy = np.ascontiguousarray(xy, np.float32)
    out = np.zeros(len(xy), np.uint64)
    load(fast).refh_coord_to_cell(nx, ny, cs, xy, len(xy), out)
    return out
