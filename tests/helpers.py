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


def force_term_magnitudes(a, r_max2=30.0, mul_repulse=1.0, mul_push=0.1, mul_scatter=0.1, max_speed=25.0):
    """Per agent: T = sum over its physics neighbours of the MAGNITUDES of the force terms applyForces adds
    (PhysicsSystem.cpp:69-116), times their multipliers. |computed sum - reference sum| of ANY implementation that
    evaluates the same terms with relative error eps is bounded by eps*T, whatever the cancellation between the terms;
    the tolerance-mode gate uses it where the plain 1e-5 gate (relative to the RESULT) is ill-posed.
    Float64 arithmetic, neighbour pairs from a KD-tree."""
    from scipy.spatial import cKDTree
    r = np.asarray(a["r"], np.float64)
    ok = np.isfinite(r).all(1)
    idx = np.nonzero(ok)[0]
    tree = cKDTree(r[idx])
    pairs = tree.query_pairs(np.sqrt(r_max2) * (1 + 1e-9), output_type="ndarray")
    i = np.concatenate([idx[pairs[:, 0]], idx[pairs[:, 1]]])
    j = np.concatenate([idx[pairs[:, 1]], idx[pairs[:, 0]]])
    dr = r[j] - r[i]
    d2 = (dr * dr).sum(1)
    keep = d2 > 0
    i, j, dr, d2 = i[keep], j[keep], dr[keep], d2[keep]
    rad = np.asarray(a["radius"], np.float64)
    mass = np.asarray(a["mass"], np.float64)
    state = np.asarray(a["state"])
    rc2 = (rad[i] + rad[j]) ** 2
    x = rc2 / d2
    w = mass[j] / (mass[i] + mass[j])
    rep = np.where(x > 0.8, np.minimum(3 * x ** 3, 50000.0) * w, 0.0) * mul_repulse
    push = np.where((state[j] == 0) & (state[i] == 1) & (x > 0.75), np.sqrt(d2) * x * w, 0.0) * mul_push
    n = len(r)
    T = np.bincount(i, rep + push, minlength=n)
    # scatter = mean_dr / |mean_dr| * max_speed: the direction of a sum of offsets; its error is the terms' error over |sum|
    mv = state[j] == 0
    sx = np.bincount(i[mv], dr[mv, 0], minlength=n)
    sy = np.bincount(i[mv], dr[mv, 1], minlength=n)
    sabs = np.bincount(i[mv], np.sqrt(d2[mv]), minlength=n)
    cnt = np.bincount(i[mv], minlength=n)
    snorm = np.hypot(sx, sy)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = np.where(cnt > 1, sabs / np.maximum(snorm, 1e-30), 0.0)
    T = T + mul_scatter * max_speed * np.minimum(amp, 1e12)
    return T
