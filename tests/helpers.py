"""Shared helpers of the test-suite: fixture loading and the agent/path layouts of the C ABI."""
from __future__ import annotations

import os

import numpy as np

from rtsengine_b200 import abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TABLE_KEYS = ["tri_verts", "tri_nbrs", "tri_constrained", "cell2tri", "n_cells"] + \
             [f"line_off{o}" for o in range(4)] + [f"edges{o}" for o in range(4)]


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def tables_of(fx, prefix=""):
    return {k: fx[prefix + k] for k in TABLE_KEYS}


def windows_to_paths(win):
    """One 2-waypoint path per agent from PathFinderComponent windows (layout of refh.seek_window):
    r_next r_nn | portal_next.from .t | portal_nn.from .t | portal_next.l portal_nn.l path_end"""
    n = len(win)
    off = (np.arange(n + 1) * 2).astype(np.uint32)
    wp = np.zeros((n, 2, 2), np.float32)
    po = np.zeros((n, 2, 5), np.float32)
    wp[:, 0] = win[:, 0:2]
    wp[:, 1] = win[:, 2:4]
    po[:, 0, 0:4] = win[:, 4:8]
    po[:, 0, 4] = win[:, 12]
    po[:, 1, 0:4] = win[:, 8:12]
    po[:, 1, 4] = win[:, 13]
    return {"offsets": off, "waypoints": wp.reshape(-1, 2), "portals": po.reshape(-1, 5),
            "path_end": np.ascontiguousarray(win[:, 14:16])}


ALL_FIELDS = [f[0] for f in abi.AGENT_FIELDS if f[0] != "gid"]


def frame_agents(fx, step, prev=None):
    """Agent arrays entering recorded frame `step` of a frame_*.npz fixture. `prev` = the arrays that
    came out of frame step-1 on the side under test (None: start from the fixture's initial state)."""
    n = len(fx["radius"])
    a = abi.alloc_agent_arrays(n, set(ALL_FIELDS))
    if prev is None:
        a["r"][:] = fx["r0"]
        a["inertia"][:] = fx["inertia0"]
        a["angle"][:] = fx["angle0"]
        a["angle_vel"][:] = fx["angle_vel0"]
    else:
        for k in ("r", "inertia", "angle", "angle_vel"):
            a[k][:] = prev[k]
    a["radius"][:] = fx["radius"]
    a["mass"][:] = fx["mass"]
    a["player"][:] = fx["player"]
    win = fx["x_win"][step]
    a["state"][:] = fx["x_state_in"][step]
    a["path_id"][:] = np.where(a["state"] == abi.MOVING, np.arange(n), -1)
    a["waypoint"][:] = 0
    a["r_next"][:] = win[:, 0:2]
    return a, windows_to_paths(win)


def assert_same_bits(got, want, what):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if got.dtype.kind == "f":
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want)) | ((got == 0) & (want == 0))
    else:
        same = got == want
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {same.size} elements differ; first at {i}: got {got[i]!r} want {want[i]!r}")


def rel_err(got, want):
    """vector-norm relative error per agent: |got-want|_2 / max(|want|_2, 1)   """
    d = np.linalg.norm(np.asarray(got, np.float64) - np.asarray(want, np.float64), axis=-1)
    return d / np.maximum(np.linalg.norm(np.asarray(want, np.float64), axis=-1), 1.0)
