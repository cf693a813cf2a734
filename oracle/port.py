"""TEST INFRASTRUCTURE — ctypes view of oracle/librts_oracle*.so (the plain-C restatement).

Only tests/, __graft_entry__.smoke(), and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rtsengine_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs: dict = {}


def lib_path(fast=False):
    return os.path.join(_HERE, "librts_oracle_fast.so" if fast else "librts_oracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def load(fast=False):
    if fast in _libs:
        return _libs[fast]
    if not os.path.exists(lib_path(fast)):
        build()
    L = C.CDLL(lib_path(fast))
    vp = C.c_void_p
    L.rts_oracle_default_params.argtypes = [C.POINTER(abi.RtsParams), C.c_int32, C.c_int32]
    L.rts_oracle_default_params.restype = None
    L.rts_oracle_step.argtypes = [C.POINTER(abi.RtsParams), vp, C.c_uint32, vp, C.POINTER(abi.RtsWallTable), vp,
                                  C.c_uint32, C.POINTER(abi.RtsPathTable), C.c_uint32, C.POINTER(abi.RtsAgentView), vp, vp]
    L.rts_oracle_step.restype = C.c_int32
    L.rts_oracle_coord_to_cell.argtypes = [vp, C.c_uint32, C.c_int32, C.c_int32, C.c_float, vp]
    L.rts_oracle_find_triangles.argtypes = [C.POINTER(abi.RtsParams), vp, C.c_uint32, vp, vp, C.c_uint32, vp]
    L.rts_oracle_contact_edges.argtypes = [C.POINTER(abi.RtsParams), C.POINTER(abi.RtsWallTable), vp, vp, C.c_uint32,
                                           C.c_uint32, vp, vp]
    L.rts_oracle_neighbours.argtypes = [C.POINTER(abi.RtsParams), C.c_uint32, vp, C.c_uint32, vp, C.c_uint32]
    L.rts_oracle_query_radius.argtypes = [C.POINTER(abi.RtsParams), C.c_uint32, vp, C.c_float, C.c_float, C.c_float, vp, C.c_uint32]
    L.rts_oracle_query_radius.restype = C.c_uint32
    L.rts_oracle_attack_targets.argtypes = [C.POINTER(abi.RtsParams), C.c_uint32, vp, vp, C.c_float, vp]
    L.rts_oracle_attack_targets.restype = None
    L.rts_oracle_neighbours.restype = C.c_uint32
    L.rts_oracle_acos_table.argtypes = [vp]
    L.rts_oracle_build_cell_grid.argtypes = [C.POINTER(abi.RtsParams), vp, C.c_uint32, C.c_uint32, vp]
    L.rts_oracle_build_cell_grid.restype = None
    L.rts_oracle_threads.restype = C.c_int
    _libs[fast] = L
    return L


def default_params(map_nx, map_ny, fast=False) -> abi.RtsParams:
    p = abi.RtsParams()
    load(fast).rts_oracle_default_params(C.byref(p), map_nx, map_ny)
    return p


class OracleWorld:
    """Holds map / path tables and steps agent arrays on the CPU (caller order, in place)."""

    def __init__(self, params: abi.RtsParams, tables: dict, paths: dict | None = None, unit_types=None, fast=False):
        self.L = load(fast)
        self.params = params
        self._keep: list = []
        self.tris = abi.triangles_array(tables["tri_verts"], tables["tri_nbrs"], tables.get("tri_constrained"))
        self.cell2tri = np.ascontiguousarray(tables["cell2tri"], np.uint32)
        self.walls = abi.wall_table(tables, self._keep)
        self.unit_types = np.ascontiguousarray(abi.default_unit_types() if unit_types is None else unit_types)
        if paths is None:
            paths = {"offsets": np.zeros(1, np.uint32), "waypoints": np.zeros((0, 2), np.float32),
                     "portals": np.zeros((0, 5), np.float32), "path_end": np.zeros((0, 2), np.float32)}
        self.paths = abi.path_table(paths, self._keep)
        self.fin_area = np.zeros(max(int(self.paths.n_groups), 1), np.float32)   # group2finished_surface_area_

    def step(self, agents: dict, n_steps=1, ghost=None):
        """agents: dict of numpy arrays (see abi.AGENT_FIELDS); updated in place."""
        n = len(agents["r"])
        keep: list = []
        for name, dt, k in abi.AGENT_FIELDS:
            if name in agents:
                a = agents[name]
                if a.dtype != dt or not a.flags.c_contiguous:
                    raise ValueError(f"{name}: need C-contiguous {dt}")
        view = abi.agent_view(agents, n, keep)
        g = None
        if ghost is not None:
            g = np.ascontiguousarray(ghost, np.uint8)
        for _ in range(n_steps):
            rc = self.L.rts_oracle_step(C.byref(self.params), self.tris.ctypes.data, len(self.tris),
                                        self.cell2tri.ctypes.data, C.byref(self.walls), self.unit_types.ctypes.data,
                                        len(self.unit_types), C.byref(self.paths), n, C.byref(view),
                                        None if g is None else g.ctypes.data, self.fin_area.ctypes.data)
            if rc != 0:
                raise RuntimeError(f"oracle step failed: {rc}")

    def set_slab_arrival(self, on: bool, counts=None):
        """slab model of the arrival rule (see rts_oracle_set_slab_arrival): ghosts evaluate the rule too; `counts` (uint32 per
        group, kept alive by the caller) accumulates the agents stopped by the following steps"""
        self._stop_counts = counts
        self.L.rts_oracle_set_slab_arrival(1 if on else 0, None if counts is None else counts.ctypes.data)

    def find_triangles(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.uint32)
        self.L.rts_oracle_find_triangles(C.byref(self.params), self.tris.ctypes.data, len(self.tris),
                                         self.cell2tri.ctypes.data, xy.ctypes.data, len(xy), out.ctypes.data)
        return out

    def build_cell_grid(self, last_found=0):
        """Triangulation::updateCellGrid restated: the cell -> triangle cache, built from the triangles alone"""
        out = np.zeros(int(self.params.tri_nx) * int(self.params.tri_ny), np.uint32)
        self.L.rts_oracle_build_cell_grid(C.byref(self.params), self.tris.ctypes.data, len(self.tris), last_found,
                                          out.ctypes.data)
        return out

    def contact_edges(self, xy, radius, cap=8):
        xy = np.ascontiguousarray(xy, np.float32)
        radius = np.ascontiguousarray(radius, np.float32)
        n = len(xy)
        counts = np.zeros(n, np.uint32)
        out = np.zeros((n, cap, 5), np.float32)
        self.L.rts_oracle_contact_edges(C.byref(self.params), C.byref(self.walls), xy.ctypes.data, radius.ctypes.data,
                                        n, cap, counts.ctypes.data, out.ctypes.data)
        return counts, out

    def query_radius(self, r, q, r_max, cap=4096):
        """NeighbourSearcherContext::getNeighboursIndsFull on the physics grid: indices within r_max of q, visit order"""
        r = np.ascontiguousarray(r, np.float32)
        out = np.zeros(cap, np.uint32)
        n = self.L.rts_oracle_query_radius(C.byref(self.params), len(r), r.ctypes.data, float(q[0]), float(q[1]), float(r_max), out.ctypes.data, cap)
        return out[:min(n, cap)].copy(), int(n)

    def attack_targets(self, r, player, range2=2560.0):
        """NeighbourSearcherStrategyAttack::execute: per agent the enemy the reference's loop ends on, or -1"""
        r = np.ascontiguousarray(r, np.float32)
        player = np.ascontiguousarray(player, np.uint8)
        out = np.zeros(len(r), np.int32)
        self.L.rts_oracle_attack_targets(C.byref(self.params), len(r), r.ctypes.data, player.ctypes.data, float(range2), out.ctypes.data)
        return out

    def neighbours(self, r, i, cap=4096):
        r = np.ascontiguousarray(r, np.float32)
        out = np.zeros(cap, np.uint32)
        k = self.L.rts_oracle_neighbours(C.byref(self.params), len(r), r.ctypes.data, i, out.ctypes.data, cap)
        return out[:min(k, cap)]


def coord_to_cell(xy, nx, ny, cs, fast=False):
    xy = np.ascontiguousarray(xy, np.float32)
    out = np.zeros(len(xy), np.uint32)
    load(fast).rts_oracle_coord_to_cell(xy.ctypes.data, len(xy), nx, ny, cs, out.ctypes.data)
    return out


def acos_table():
    t = np.zeros(1000, np.float32)
    load().rts_oracle_acos_table(t.ctypes.data)
    return t
