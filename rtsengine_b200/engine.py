"""Host-side binding of librts_b200.so (C ABI in include/rts_gpu.h).

`GpuAgentSystem` mirrors the three-phase interface of the reference's `System2`
(src/ECS.h:147-234: updateSharedData -> update -> communicate) for the fused Physics+Seek path.
There is no CPU fallback: if the CUDA library cannot be loaded or no device is present this
module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTS_B200_LIB") or os.path.join(_HERE, "librts_b200.so")   # override: A/B of build variants (tools/)

_lib = None


class RtsError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"{abi.STATUS_NAMES.get(code, code)}: {text}")
        self.code = code


def load_library():
    """Loads librts_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C rtsengine_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    i32, u32, f32 = C.c_int32, C.c_uint32, C.c_float
    vp = C.c_void_p
    PP = C.POINTER(abi.RtsParams)
    AV = C.POINTER(abi.RtsAgentView)
    sig = {
        "rts_default_params": (None, [PP, i32, i32]),
        "rts_create": (i32, [PP, i32, C.POINTER(H)]),
        "rts_destroy": (i32, [H]),
        "rts_last_error": (C.c_char_p, [H]),
        "rts_set_params": (i32, [H, PP]),
        "rts_upload_map": (i32, [H, vp, u32, vp, C.POINTER(abi.RtsWallTable)]),
        "rts_update_map_region": (i32, [H, u32, vp, vp, u32, vp, vp, vp, vp, u32, u32]),
        "rts_map_upload_bytes": (i32, [H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "rts_get_cell_grid": (i32, [H, vp]),
        "rts_upload_unit_types": (i32, [H, vp, u32]),
        "rts_upload_paths": (i32, [H, C.POINTER(abi.RtsPathTable)]),
        "rts_set_agents": (i32, [H, u32, AV]),
        "rts_append_agents": (i32, [H, u32, AV]),
        "rts_remove_agents": (i32, [H, vp, u32]),
        "rts_command_agents": (i32, [H, vp, u32, i32, C.c_uint8]),
        "rts_update_agents": (i32, [H, AV]),
        "rts_update_agents_indexed": (i32, [H, vp, u32, AV]),
        "rts_agent_count": (u32, [H]),
        "rts_get_finished_area": (i32, [H, vp, u32]),
        "rts_set_finished_area": (i32, [H, vp, u32]),
        "rts_step": (i32, [H, u32]),
        "rts_download": (i32, [H, AV]),
        "rts_download_delta": (i32, [H, vp, vp, u32, C.POINTER(C.c_uint32), vp, vp, vp]),
        "rts_apply_delta": (i32, [vp, u32, u32, vp, vp, vp]),
        "rts_sync": (i32, [H]),
        "rts_get_timings": (i32, [H, C.POINTER(abi.RtsTimings)]),
        "rts_kernel_time_ms": (i32, [H, i32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
        "rts_reset_timings": (i32, [H]),
        "rts_set_profiling": (i32, [H, i32]),
        "rts_abi_sizeof": (C.c_size_t, [i32]),
        "rts_coord_to_cell": (i32, [H, vp, u32, i32, i32, f32, vp]),
        "rts_find_triangles": (i32, [H, vp, u32, vp]),
        "rts_build_cell_grid": (i32, [H, u32, vp]),
        "rts_contact_edges": (i32, [H, vp, vp, u32, u32, vp, vp]),
        "rts_stream": (vp, [H]),
        "rts_query_radius": (i32, [H, vp, vp, u32, u32, vp, vp]),
        "rts_attack_targets": (i32, [H, f32, vp]),
        "rts_slab_configure": (i32, [H, i32, i32, u32, u32]),
        "rts_comm_unique_id": (i32, [vp]),
        "rts_comm_init": (i32, [H, vp]),
        "rts_slab_exchange_mode": (i32, [H]),
        "rts_slab_row_counts": (i32, [H, vp]),
        "rts_slab_set_rows": (i32, [H, i32, i32]),
        "rts_halo_refresh": (i32, [H]),
        "rts_step_begin": (i32, [H]),
        "rts_halo_begin": (i32, [H]),
        "rts_step_end": (i32, [H]),
        "rts_slab_exchange_local": (i32, [H, H]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._rts_signatures = sig
    _lib = L
    return L


def default_params(map_nx: int, map_ny: int) -> abi.RtsParams:
    p = abi.RtsParams()
    load_library().rts_default_params(C.byref(p), map_nx, map_ny)
    return p


KERNELS = {"step": 0, "bin": 1, "scan": 2, "reorder": 3, "exchange": 4, "ingest": 5, "walk": 6, "dense": 7}


class GpuAgentSystem:
    """The fused Physics+Seek system on one GPU.

    reference interface                         here
    ------------------------------------------  -----------------------------------------------
    System2::addComponent (ECS.h:165)           set_agents / append_agents
    System2::remove (ECS.h:197)                 remove_agents
    System2::updateSharedData (ECS.h:232)       update_shared_data(fields)
    System2::update (ECS.h:231)                 update(n_steps)
    System2::communicate (ECS.h:233)            communicate(names) -> dict of arrays
    PhysicsSystem::p_map_, SeekSystem::p_cdt_   upload_map(tables)
    PathFinder2::setPathOfAgent                 upload_paths(paths)
    PhysicsSystem::force_multipliers            params / set_params
    """

    def __init__(self, params: abi.RtsParams | None = None, map_cells=(512, 512), device: int = 0):
        self.L = load_library()
        self.params = params if params is not None else default_params(*map_cells)
        self.h = C.c_void_p()
        rc = self.L.rts_create(C.byref(self.params), device, C.byref(self.h))
        if rc != 0:
            text = self.L.rts_last_error(self.h).decode() if self.h else "no CUDA device (there is no CPU fallback)"
            if self.h:
                self.L.rts_destroy(self.h)
                self.h = C.c_void_p()
            raise RtsError(rc, text)
        self._keep: list = []

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.rts_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            raise RtsError(rc, self.L.rts_last_error(self.h).decode())

    # -- tables -----------------------------------------------------------------------------
    def set_params(self, params: abi.RtsParams):
        self._chk(self.L.rts_set_params(self.h, C.byref(params)))
        self.params = params

    def upload_map(self, tables: dict):
        keep: list = []
        tris = abi.triangles_array(tables["tri_verts"], tables["tri_nbrs"], tables.get("tri_constrained"))
        c2t = tables.get("cell2tri")   # None: the cache is built on the device (Triangulation::updateCellGrid)
        if c2t is not None:
            c2t = np.ascontiguousarray(c2t, np.uint32)
            if c2t.size != self.params.tri_nx * self.params.tri_ny:
                raise ValueError("cell2tri does not match the triangle grid of the params")
        walls = abi.wall_table(tables, keep)
        self._chk(self.L.rts_upload_map(self.h, tris.ctypes.data, len(tris), None if c2t is None else c2t.ctypes.data,
                                        C.byref(walls)))

    def update_map_region(self, before: dict, after: dict, last_found: int = 0):
        """Game::updateTriangulation after a building insert: diff of two table sets (as Triangulation::triangles_ and the
        edge finders hold them before / after the insert) -> rts_update_map_region. Returns (changed triangles, changed lines)."""
        tb = abi.triangles_array(before["tri_verts"], before["tri_nbrs"], before.get("tri_constrained"))
        ta = abi.triangles_array(after["tri_verts"], after["tri_nbrs"], after.get("tri_constrained"))
        nb, na = len(tb), len(ta)
        same = np.zeros(na, bool)
        same[:nb] = (tb.view(np.uint8).reshape(nb, -1) == ta[:nb].view(np.uint8).reshape(nb, -1)).all(1)
        idx = np.nonzero(~same)[0].astype(np.uint32)
        recs = np.ascontiguousarray(ta[idx])
        orient, index, counts, edges = [], [], [], []
        for o in range(4):
            ob, oa = before[f"line_off{o}"], after[f"line_off{o}"]
            eb, ea = before[f"edges{o}"], after[f"edges{o}"]
            for l in range(len(oa) - 1):
                new = ea[oa[l]:oa[l + 1]]
                old = eb[ob[l]:ob[l + 1]]
                if new.shape != old.shape or not np.array_equal(new.view(np.uint32), old.view(np.uint32)):
                    orient.append(o); index.append(l); counts.append(len(new)); edges.append(new)
        e = np.ascontiguousarray(np.concatenate(edges).astype(np.float32)) if edges else np.zeros((0, 5), np.float32)
        orient, index, counts = (np.array(v, np.uint32) for v in (orient, index, counts))
        self._chk(self.L.rts_update_map_region(self.h, na, idx.ctypes.data, recs.ctypes.data, len(idx), orient.ctypes.data, index.ctypes.data,
                                               counts.ctypes.data, e.ctypes.data, len(orient), int(last_found)))
        return len(idx), len(orient)

    def map_upload_bytes(self):
        a, b = C.c_uint64(), C.c_uint64()
        self._chk(self.L.rts_map_upload_bytes(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def cell_grid(self) -> np.ndarray:
        out = np.zeros(int(self.params.tri_nx) * int(self.params.tri_ny), np.uint32)
        self._chk(self.L.rts_get_cell_grid(self.h, out.ctypes.data))
        return out

    def build_cell_grid(self, last_found=0) -> np.ndarray:
        """Triangulation::updateCellGrid on the device; returns (and installs) the cell -> triangle cache"""
        out = np.zeros(int(self.params.tri_nx) * int(self.params.tri_ny), np.uint32)
        self._chk(self.L.rts_build_cell_grid(self.h, int(last_found), out.ctypes.data))
        return out

    def upload_unit_types(self, types: np.ndarray):
        types = np.ascontiguousarray(types, abi.UNIT_DTYPE)
        self._chk(self.L.rts_upload_unit_types(self.h, types.ctypes.data, len(types)))

    def upload_paths(self, paths: dict):
        keep: list = []
        t = abi.path_table(paths, keep)
        self._chk(self.L.rts_upload_paths(self.h, C.byref(t)))

    # -- agents -----------------------------------------------------------------------------
    def set_agents(self, agents: dict):
        n = len(agents["r"])
        keep: list = []
        v = abi.agent_view({k: a for k, a in agents.items() if k not in abi.DOWNLOAD_ONLY}, n, keep)
        self._chk(self.L.rts_set_agents(self.h, n, C.byref(v)))

    def append_agents(self, agents: dict):
        n = len(agents["r"])
        keep: list = []
        v = abi.agent_view({k: a for k, a in agents.items() if k not in abi.DOWNLOAD_ONLY}, n, keep)
        self._chk(self.L.rts_append_agents(self.h, n, C.byref(v)))

    def remove_agents(self, indices):
        idx = np.ascontiguousarray(indices, np.uint32)
        self._chk(self.L.rts_remove_agents(self.h, idx.ctypes.data, len(idx)))

    def command_agents(self, indices, path_id: int, state: int = abi.MOVING):
        idx = np.ascontiguousarray(indices, np.uint32)
        self._chk(self.L.rts_command_agents(self.h, idx.ctypes.data, len(idx), path_id, state))

    def finished_area(self, n_groups: int) -> np.ndarray:
        out = np.zeros(n_groups, np.float32)
        self._chk(self.L.rts_get_finished_area(self.h, out.ctypes.data, n_groups))
        return out

    def set_finished_area(self, area):
        """the groups' finished areas (CommandGroupManager2::group2finished_surface_area_) — with the agent arrays, the whole
        persistent state of the path: a run continues from a download on another handle (tests: snapshot and resume)"""
        a = np.ascontiguousarray(area, np.float32)
        self._chk(self.L.rts_set_finished_area(self.h, a.ctypes.data, len(a)))

    @property
    def n(self) -> int:
        return self.L.rts_agent_count(self.h)

    # -- the three System2 phases -------------------------------------------------------------
    def update_shared_data(self, fields: dict):
        keep: list = []
        v = abi.agent_view({k: a for k, a in fields.items() if k not in abi.DOWNLOAD_ONLY}, self.n, keep)
        self._chk(self.L.rts_update_agents(self.h, C.byref(v)))

    def update_shared_data_indexed(self, indices, fields: dict):
        """System2::updateSharedData for the entities the host touched: `fields` arrays hold len(indices) entries"""
        idx = np.ascontiguousarray(indices, np.uint32)
        keep: list = []
        v = abi.agent_view({k: a for k, a in fields.items() if k not in abi.DOWNLOAD_ONLY}, len(idx), keep)
        self._chk(self.L.rts_update_agents_indexed(self.h, idx.ctypes.data, len(idx), C.byref(v)))

    def update_shared_data_indexed_ptrs(self, idx_ptr: int, m: int, ptrs: dict):
        v = abi.RtsAgentView()
        for k, p in ptrs.items():
            setattr(v, k, p)
        self._chk(self.L.rts_update_agents_indexed(self.h, idx_ptr, m, C.byref(v)))

    def update(self, n_steps: int = 1, sync: bool = False):
        self._chk(self.L.rts_step(self.h, n_steps))
        if sync:
            self.sync()

    step = update

    def communicate(self, names=("r", "vel", "inertia", "angle", "angle_vel", "state"), out: dict | None = None) -> dict:
        self.sync()   # a slab handle's population (owned + ghosts) changes every frame: refresh the count first
        n = self.n
        arrays = out if out is not None else abi.alloc_agent_arrays(n, set(names))
        keep: list = []
        dtypes = {name: dt for name, dt, _ in abi.AGENT_FIELDS}
        for k in names:
            if arrays[k].dtype != dtypes[k] or not arrays[k].flags.c_contiguous:
                raise ValueError(f"output array {k!r} must be C-contiguous with the ABI dtype")
        v = abi.agent_view({k: arrays[k] for k in names}, n, keep)
        self._chk(self.L.rts_download(self.h, C.byref(v)))
        return arrays

    download = communicate

    def communicate_delta(self, board: dict, changes: np.ndarray | None = None) -> int:
        """System2::communicate as a difference (rts_download_delta): `board` holds the blackboard of the previous frame —
        vel (n,2) f32, angle_vel (n,) f32, state (n,) u8, r_next (n,2) f32 — and is brought up to date in place; returns the
        number of entities whose angle_vel / state / target travelled."""
        n = self.n
        if changes is None:
            changes = np.zeros(n, abi.DELTA_DTYPE)
        cnt = C.c_uint32(0)
        self._chk(self.L.rts_download_delta(self.h, board["vel"].ctypes.data, changes.ctypes.data, len(changes), C.byref(cnt),
                                            board["angle_vel"].ctypes.data, board["state"].ctypes.data, board["r_next"].ctypes.data))
        return int(cnt.value)

    def sync(self):
        self._chk(self.L.rts_sync(self.h))

    # -- raw-pointer variants for pinned host memory (bench end-to-end leg) -------------------------
    def update_shared_data_ptrs(self, ptrs: dict):
        v = abi.RtsAgentView()
        for k, p in ptrs.items():
            setattr(v, k, p)
        self._chk(self.L.rts_update_agents(self.h, C.byref(v)))

    def communicate_ptrs(self, ptrs: dict):
        v = abi.RtsAgentView()
        for k, p in ptrs.items():
            setattr(v, k, p)
        self._chk(self.L.rts_download(self.h, C.byref(v)))

    # -- measurement ------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._chk(self.L.rts_set_profiling(self.h, 1 if on else 0))

    def reset_timings(self):
        self._chk(self.L.rts_reset_timings(self.h))

    def timings(self) -> abi.RtsTimings:
        t = abi.RtsTimings()
        self._chk(self.L.rts_get_timings(self.h, C.byref(t)))
        return t

    def kernel_time_ms(self, which="step"):
        ms = C.c_double()
        cnt = C.c_uint64()
        self._chk(self.L.rts_kernel_time_ms(self.h, KERNELS[which], C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # -- stand-alone queries ------------------------------------------------------------------
    def coord_to_cell(self, xy, nx, ny, cs):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.uint32)
        self._chk(self.L.rts_coord_to_cell(self.h, xy.ctypes.data, len(xy), nx, ny, cs, out.ctypes.data))
        return out

    def find_triangles(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        out = np.zeros(len(xy), np.uint32)
        self._chk(self.L.rts_find_triangles(self.h, xy.ctypes.data, len(xy), out.ctypes.data))
        return out

    def query_radius(self, xy, r_max, cap=256):
        """NeighbourSearcherContext::getNeighboursIndsFull for every query point: (counts, indices[n, cap])"""
        xy = np.ascontiguousarray(xy, np.float32)
        r_max = np.ascontiguousarray(r_max, np.float32)
        counts = np.zeros(len(xy), np.uint32)
        out = np.zeros((len(xy), cap), np.uint32)
        self._chk(self.L.rts_query_radius(self.h, xy.ctypes.data, r_max.ctypes.data, len(xy), cap, counts.ctypes.data, out.ctypes.data))
        return counts, out

    def attack_targets(self, range2=2560.0):
        """NeighbourSearcherStrategyAttack::execute for every agent: the enemy index its loop ends on, or -1"""
        out = np.zeros(self.n, np.int32)
        self._chk(self.L.rts_attack_targets(self.h, float(range2), out.ctypes.data))
        return out

    def contact_edges(self, xy, radius, cap=8):
        xy = np.ascontiguousarray(xy, np.float32)
        radius = np.ascontiguousarray(radius, np.float32)
        n = len(xy)
        counts = np.zeros(n, np.uint32)
        out = np.zeros((n, cap, 5), np.float32)
        self._chk(self.L.rts_contact_edges(self.h, xy.ctypes.data, radius.ctypes.data, n, cap, counts.ctypes.data,
                                           out.ctypes.data))
        return counts, out
