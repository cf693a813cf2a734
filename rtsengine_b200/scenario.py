"""Synthetic populations for the BASELINE.json configurations.

Maps (CDT, wall tables, per-group funnel paths) are inputs of the hot path: they are built on the
host by the reference's own Triangulation / PathFinder2 and committed as fixtures under
tests/golden/ (tests/golden/make_golden.py).  This module only places agents.
"""
from __future__ import annotations

import os

import numpy as np

from . import abi

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TABLE_KEYS = ["tri_verts", "tri_nbrs", "tri_constrained", "cell2tri", "n_cells"] + \
             [f"line_off{o}" for o in range(4)] + [f"edges{o}" for o in range(4)]


def load_map(name: str) -> dict:
    return dict(np.load(os.path.join(GOLDEN, name)))


def tables_of(fx: dict) -> dict:
    return {k: fx[k] for k in TABLE_KEYS}


def paths_of(fx: dict) -> dict:
    return {"offsets": fx["path_offsets"], "waypoints": fx["path_waypoints"], "portals": fx["path_portals"],
            "path_end": fx["path_end"]}


def occupancy(fx: dict, margin_cells: int = 1) -> np.ndarray:
    """MapGrid cells covered by buildings (+margin), [x, y]"""
    nx, ny = (int(v) for v in fx["n_cells"])
    occ = np.zeros((nx, ny), bool)
    for cx, cy, n, m in fx["buildings"]:
        x0, y0 = cx - n // 2, cy - m // 2
        occ[max(x0 - margin_cells, 0):x0 + n + margin_cells, max(y0 - margin_cells, 0):y0 + m + margin_cells] = True
    return occ


def free_positions(fx: dict, n: int, rng: np.random.Generator, lo=None, hi=None) -> np.ndarray:
    """uniform over free space by rejection with a one-cell margin (radius+1 <= 6 units for the largest unit type)"""
    nx, ny = (int(v) for v in fx["n_cells"])
    W, Hh = nx * 5.0, ny * 5.0
    lo = np.array([8.0, 8.0]) if lo is None else np.asarray(lo, float)
    hi = np.array([W - 8.0, Hh - 8.0]) if hi is None else np.asarray(hi, float)
    occ = occupancy(fx)
    out = np.empty((n, 2), np.float32)
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.1) + 64
        p = rng.uniform(lo, hi, (m, 2)).astype(np.float32)
        ok = ~occ[(p[:, 0] // 5).astype(np.int64), (p[:, 1] // 5).astype(np.int64)]
        p = p[ok][: n - filled]
        out[filled:filled + len(p)] = p
        filled += len(p)
    return out


def _blank(n: int) -> dict:
    names = {f[0] for f in abi.AGENT_FIELDS if f[0] not in abi.DOWNLOAD_ONLY and f[0] != "gid"}
    a = abi.alloc_agent_arrays(n, names)
    a["radius"][:] = 3.0
    a["mass"][:] = 1.0
    a["state"][:] = abi.STANDING
    a["path_id"][:] = -1
    a["r_next"][:] = np.nan
    return a


def config3_agents(fx: dict, n: int = 1_000_000, seed: int = 1236):
    """BASELINE config 3: n agents over the whole 1024² map, radius 3 / mass 1, all MOVING, one command group per
    block of the 8x8 block grid, each following its group's precomputed funnel path (waypoint 0 is the path's
    start point, so agents begin at waypoint 1 as PathFinder2::setPathOfAgent does, PathFinder2.cpp:338)."""
    rng = np.random.default_rng(seed)
    a = _blank(n)
    a["r"][:] = free_positions(fx, n, rng)
    a["angle"][:] = rng.uniform(-180, 180, n)
    nx = int(fx["n_cells"][0])
    n_groups = len(fx["path_end"])
    side = int(round(np.sqrt(n_groups)))
    block = nx * 5.0 / side
    g = np.minimum((a["r"][:, 0] // block).astype(np.int64), side - 1) + side * np.minimum((a["r"][:, 1] // block).astype(np.int64), side - 1)
    a["path_id"][:] = g
    a["waypoint"][:] = 1
    a["state"][:] = abi.MOVING
    return a, paths_of(fx)


def config2_agents(fx: dict, n: int = 100_000, seed: int = 1235):
    """BASELINE config 2: n agents, 80 % "default" (RADIUS 3, MASS 1) / 20 % "new unit" (RADIUS 5, MASS 50) from
    UnitTypes.txt, 50 % MOVING in 16 command groups."""
    rng = np.random.default_rng(seed)
    a = _blank(n)
    a["r"][:] = free_positions(fx, n, rng)
    a["angle"][:] = rng.uniform(-180, 180, n)
    big = rng.random(n) < 0.2
    a["radius"][:] = np.where(big, 5.0, 3.0)
    a["mass"][:] = np.where(big, 50.0, 1.0)
    a["unit_type"][:] = big
    a["player"][:] = rng.integers(0, 2, n)
    moving = rng.random(n) < 0.5
    a["state"][:] = np.where(moving, abi.MOVING, abi.STANDING)
    a["path_id"][:] = np.where(moving, rng.integers(0, 16, n) * 4, -1)
    a["waypoint"][:] = 1
    return a, paths_of(fx)


def config1_agents(n: int = 2000, seed: int = 1234):
    """BASELINE config 1: n agents in [200,1200]² of an empty 512² map, all MOVING to (2000,2000) in one group."""
    rng = np.random.default_rng(seed)
    a = _blank(n)
    a["r"][:] = rng.uniform(200, 1200, (n, 2))
    a["state"][:] = abi.MOVING
    a["path_id"][:] = 0
    a["waypoint"][:] = 0
    paths = {"offsets": np.array([0, 1], np.uint32), "waypoints": np.array([[2000.0, 2000.0]], np.float32),
             "portals": np.zeros((1, 5), np.float32), "path_end": np.array([[2000.0, 2000.0]], np.float32)}
    return a, paths
