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


def free_positions(fx: dict, n: int, rng: np.random.Generator, lo=None, hi=None, margin_cells: int = 1) -> np.ndarray:
    """uniform over free space by rejection with a one-cell margin (radius+1 <= 6 units for the largest unit type)"""
    nx, ny = (int(v) for v in fx["n_cells"])
    W, Hh = nx * 5.0, ny * 5.0
    lo = np.array([8.0, 8.0]) if lo is None else np.asarray(lo, float)
    hi = np.array([W - 8.0, Hh - 8.0]) if hi is None else np.asarray(hi, float)
    occ = occupancy(fx, margin_cells)
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


def path_of_position(fx: dict, r: np.ndarray) -> np.ndarray:
    """index of the funnel path an agent at r follows: its command group is the block of the side x side block grid it
    stands in, and inside the group the path planned from its sub-block (one path per start region, as
    PathFinder2::issuePaths2 buckets a group by start triangle, PathFinder2.cpp:541-606)"""
    nx = int(fx["n_cells"][0]) if "tile_cells" not in fx else int(fx["tile_cells"][0])
    side, sub = (int(v) for v in fx["path_grid"])
    W = nx * 5.0
    bw = W / side
    sw = bw / sub
    bx = np.minimum((r[:, 0] // bw).astype(np.int64), side - 1)
    by = np.minimum((r[:, 1] // bw).astype(np.int64), side - 1)
    si = np.minimum(((r[:, 0] - bx * bw) // sw).astype(np.int64), sub - 1)
    sj = np.minimum(((r[:, 1] - by * bw) // sw).astype(np.int64), sub - 1)
    return ((bx + side * by) * sub * sub + sj * sub + si).astype(np.int32)


def config3_agents(fx: dict, n: int = 1_000_000, seed: int = 1236):
    """BASELINE config 3: n agents over the whole 1024² map, radius 3 / mass 1, all MOVING in 64 command groups (8x8
    blocks, one destination each) that follow precomputed funnel paths (waypoint 0 is the path's start point, so
    agents begin at waypoint 1 as PathFinder2::setPathOfAgent does, PathFinder2.cpp:338)."""
    rng = np.random.default_rng(seed)
    a = _blank(n)
    a["r"][:] = free_positions(fx, n, rng)
    a["angle"][:] = rng.uniform(-180, 180, n)
    a["path_id"][:] = path_of_position(fx, a["r"])
    a["waypoint"][:] = 1
    a["state"][:] = abi.MOVING
    return a, paths_of(fx)


def config2_agents(fx: dict, n: int = 100_000, seed: int = 1235):
    """BASELINE config 2: n agents, 80 % "default" (RADIUS 3, MASS 1) / 20 % "new unit" (RADIUS 5, MASS 50) from
    UnitTypes.txt, 50 % MOVING in 16 command groups."""
    rng = np.random.default_rng(seed)
    a = _blank(n)
    a["r"][:] = free_positions(fx, n, rng, margin_cells=0)      # right up to the walls: some agents start in contact
    a["angle"][:] = rng.uniform(-180, 180, n)
    big = rng.random(n) < 0.2
    a["radius"][:] = np.where(big, 5.0, 3.0)
    a["mass"][:] = np.where(big, 50.0, 1.0)
    a["unit_type"][:] = big
    a["player"][:] = rng.integers(0, 2, n)
    moving = rng.random(n) < 0.5
    a["state"][:] = np.where(moving, abi.MOVING, abi.STANDING)
    n_paths = len(fx["path_end"])
    sub2 = int(fx["path_grid"][1]) ** 2
    # 16 of the 64 destinations; an agent takes the path of that group planned from a random start region
    a["path_id"][:] = np.where(moving, (rng.integers(0, 16, n) * 4) * sub2 + rng.integers(0, sub2, n), -1)
    assert a["path_id"].max() < n_paths
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


# ---------------------------------------------------------------------------------------------------------
# Large maps: tx x ty copies of a tile map.  Each tile keeps its complete reference-built CDT (super-triangle
# included), shifted to the tile's origin; the cell->triangle cache is tile aligned (5120 = 128 x 40), so
# Triangulation::findTriangle for a point of tile T starts in T's table and never leaves it.  The wall
# tables are rebuilt for the big MapGrid with the reference's line-index rules (EdgesFinder1D.cpp:231-236) and
# per-line sort keys (MapGrid.cpp:1277-1303).
# ---------------------------------------------------------------------------------------------------------
def tile_map(fx: dict, tx: int, ty: int) -> dict:
    nx_t, ny_t = (int(v) for v in fx["n_cells"])
    nx, ny = nx_t * tx, ny_t * ty
    W_t, H_t = nx_t * 5, ny_t * 5
    n_tri = len(fx["tri_verts"])
    tgx, tgy = nx_t // 8, ny_t // 8                      # triangle cache cells per tile
    # Global triangle order: the inner triangles (all vertices inside the tile square) of every tile first, then the
    # outer ones (super-triangle region).  findTriangle's rare linear-scan fallback returns the lowest index that
    # contains the query; with this order it finds the triangle of the query's own tile, as on the single tile.
    tv = fx["tri_verts"].reshape(n_tri, 3, 2)
    inner = ((tv[:, :, 0] >= 0) & (tv[:, :, 0] <= W_t) & (tv[:, :, 1] >= 0) & (tv[:, :, 1] <= H_t)).all(1)
    n_in = int(inner.sum())
    n_tiles = tx * ty
    rank_in = np.cumsum(inner) - 1
    rank_out = np.cumsum(~inner) - 1

    def gindex(t):   # local triangle index -> global index, tile t
        return np.where(inner, t * n_in + rank_in, n_tiles * n_in + t * (n_tri - n_in) + rank_out).astype(np.int64)

    tri_global = np.zeros((n_tiles * n_tri, 6), np.int32)
    nbr_global = np.zeros((n_tiles * n_tri, 3), np.uint32)
    con_global = np.zeros((n_tiles * n_tri, 3), np.uint8)
    verts, nbrs, cons = [], [], []
    c2t = np.zeros((tgy * ty, tgx * tx), np.uint32)      # [row(y), col(x)]: cell = x + y * (tgx*tx)
    tile_c2t = fx["cell2tri"].reshape(tgy, tgx)
    edges = {o: [] for o in range(4)}
    lines = {o: [] for o in range(4)}
    blds = []
    poff, pw, pp, pe, ps = [0], [], [], [], []
    for j in range(ty):
        for i in range(tx):
            t = j * tx + i
            off = np.array([i * W_t, j * H_t], np.int32)
            gi = gindex(t)
            tri_global[gi] = (tv + off).reshape(n_tri, 6)
            nb = fx["tri_nbrs"].astype(np.int64)
            nbr_global[gi] = np.where(nb == 0xFFFFFFFF, 0xFFFFFFFF, gi[np.minimum(nb, n_tri - 1)]).astype(np.uint32)
            con_global[gi] = fx["tri_constrained"]
            blk = tile_c2t.astype(np.int64)
            c2t[j * tgy:(j + 1) * tgy, i * tgx:(i + 1) * tgx] = np.where(blk == 0xFFFFFFFF, 0xFFFFFFFF, gi[np.minimum(blk, n_tri - 1)]).astype(np.uint32)
            ox, oy = i * nx_t, j * ny_t
            for o in range(4):
                e = fx[f"edges{o}"].copy()
                lo = fx[f"line_off{o}"]
                local_line = np.repeat(np.arange(len(lo) - 1), np.diff(lo))
                e[:, 0] += off[0]
                e[:, 1] += off[1]
                if o == 0:
                    gl = local_line + oy
                elif o == 2:
                    gl = local_line + ox
                elif o == 1:
                    gl = local_line - (ny_t - 1) + ox - oy + (ny - 1)
                else:
                    gl = local_line + ox + oy
                edges[o].append(e)
                lines[o].append(gl)
            b = fx["buildings"].copy()
            b[:, 0] += ox
            b[:, 1] += oy
            blds.append(b)
            if "path_offsets" in fx:
                po = fx["path_offsets"]
                poff.extend((po[1:].astype(np.int64) + poff[-1]).tolist())
                pw.append(fx["path_waypoints"] + off.astype(np.float32))
                q = fx["path_portals"].copy()
                q[:, 0] += off[0]
                q[:, 1] += off[1]
                pp.append(q)
                pe.append(fx["path_end"] + off.astype(np.float32))
                ps.append(fx["path_start"] + off.astype(np.float32))
    out = {"tri_verts": tri_global, "tri_nbrs": nbr_global,
           "tri_constrained": con_global, "cell2tri": c2t.reshape(-1), "n_cells": np.array([nx, ny], np.int32),
           "tile_tri_index": np.stack([gindex(t) for t in range(n_tiles)]).astype(np.uint32),
           "buildings": np.concatenate(blds), "tiles": np.array([tx, ty], np.int32), "tile_cells": np.array([nx_t, ny_t], np.int32)}
    n_lines = [ny + 1, nx + ny + 1, nx + 1, nx + ny + 1]
    for o in range(4):
        e = np.concatenate(edges[o])
        gl = np.concatenate(lines[o])
        if o == 0:
            key = np.minimum(e[:, 0], e[:, 0] + e[:, 2] * e[:, 4])   # min(from.x, to().x)
        elif o == 2:
            key = e[:, 1]                                            # from.y
        else:
            key = e[:, 0]                                            # from.x
        order = np.lexsort((key, gl))
        e, gl = e[order], gl[order]
        cnt = np.bincount(gl, minlength=n_lines[o])
        out[f"edges{o}"] = np.ascontiguousarray(e, np.float32)
        out[f"line_off{o}"] = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    if pw:
        out.update(path_offsets=np.array(poff, np.uint32), path_waypoints=np.concatenate(pw).astype(np.float32),
                   path_portals=np.concatenate(pp).astype(np.float32), path_end=np.concatenate(pe).astype(np.float32),
                   path_start=np.concatenate(ps).astype(np.float32), path_grid=fx["path_grid"])
    return out


def tiled_agents(fx_tile: dict, tx: int, ty: int, n_per_tile: int, seed: int = 1237):
    """config-3 population on every tile of a tile_map(fx_tile, tx, ty): n_per_tile agents per tile, all MOVING along
    the funnel path of their start region inside their own tile."""
    nx_t = int(fx_tile["n_cells"][0])
    W_t = nx_t * 5.0
    n_paths_tile = len(fx_tile["path_end"])
    parts = []
    for j in range(ty):
        for i in range(tx):
            t = j * tx + i
            rng = np.random.default_rng(seed + 7919 * t)
            a = _blank(n_per_tile)
            r = free_positions(fx_tile, n_per_tile, rng)
            a["angle"][:] = rng.uniform(-180, 180, n_per_tile)
            a["path_id"][:] = t * n_paths_tile + path_of_position(fx_tile, r)
            a["waypoint"][:] = 1
            a["state"][:] = abi.MOVING
            a["r"][:] = r + np.array([i * W_t, j * W_t], np.float32)
            parts.append(a)
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 5: dense crowd funnelled through narrow corridors (tests/golden/map_c5.npz, built through the
# reference's CDT / Edges code by tests/golden/make_golden.py::make_map_c5)
# ---------------------------------------------------------------------------------------------------------
def config5_agents(fx: dict, n: int = 4_000_000, seed: int = 1238, lo=None, hi=None):
    """n agents uniform over the free space of the window [lo, hi] (default: the whole map — 4M agents are 0.16 per
    unit², four times config 3), radius 3 / mass 1, all MOVING: the left half heads right through the corridor nearest
    to it, the right half heads left (counter-flow in every corridor). Paths: mouth -> far mouth -> plaza."""
    rng = np.random.default_rng(seed)
    n_corr, pitch, width, x0, n_seg, seg, gap = (int(v) for v in fx["corridors"])
    W = int(fx["n_cells"][0]) * 5.0
    xa, xb = x0 * 5.0, (x0 + n_seg * (seg + gap) - gap) * 5.0
    a = _blank(n)
    a["r"][:] = free_positions(fx, n, rng, lo=lo, hi=hi)
    a["angle"][:] = rng.uniform(-180, 180, n)
    x, y = a["r"][:, 0], a["r"][:, 1]
    k = np.clip(np.rint((y / 5.0 - width / 2.0) / pitch), 0, n_corr - 1).astype(np.int64)
    d = (x >= W / 2).astype(np.int64)
    a["path_id"][:] = 2 * k + d
    before = np.where(d == 0, x < xa, x > xb)
    inside = (x >= xa) & (x <= xb)
    a["waypoint"][:] = np.where(before, 1, np.where(inside, 2, 3))
    a["state"][:] = abi.MOVING
    return a, paths_of(fx)
