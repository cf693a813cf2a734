"""Generates the golden fixtures under tests/golden/ by running the VERBATIM reference.

Run here (the container that has /root/reference), after `make -C oracle ref`:

    python tests/golden/make_golden.py

Every expected value in the .npz files comes out of oracle/_ref/libref_harness.so, i.e. the
reference's own translation units compiled in place with IEEE flags (-O2 -ffp-contract=off);
nothing is computed by the oracle port or by the CUDA path.  The fixtures travel to the GPU box,
/root/reference does not.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def place_buildings(w, n_buildings, rng, lo, hi, cells):
    """random non-touching octagonal buildings (types 8x8 and 4x6, corner 1; BuildingManager.h:19-20)"""
    occupied = np.zeros(cells, bool)
    out = []
    tries = 0
    while len(out) < n_buildings and tries < 100 * n_buildings:
        tries += 1
        cx, cy = (int(v) for v in rng.integers(lo, hi, 2))
        n, m = (8, 8) if rng.random() < 0.5 else (4, 6)
        x0, y0 = cx - n // 2, cy - m // 2
        if occupied[max(x0 - 2, 0):x0 + n + 2, max(y0 - 2, 0):y0 + m + 2].any():
            continue
        if w.add_building(cx, cy, n, m, 1) != 0:
            raise RuntimeError(w.L.refh_last_error(w.h).decode())
        occupied[x0 - 1:x0 + n + 1, y0 - 1:y0 + m + 1] = True
        out.append((cx, cy, n, m))
    return np.array(out, np.int32), occupied


def spawn(rng, n, lo, hi, occupied, min_sep=0.0):
    pos = []
    while len(pos) < n:
        p = rng.uniform(lo, hi, 2).astype(np.float32)
        if occupied[int(p[0] // 5), int(p[1] // 5)]:
            continue
        pos.append(p)
    return np.array(pos, np.float32)


def record_steps(w, n_steps, extra=None):
    """windows before each frame (after the planner ran) + blackboard after it"""
    rec = {"win": [], "state_in": [], "r": [], "vel": [], "inertia": [], "angle": [], "angle_vel": [], "state": [],
           "tri": []}
    for _ in range(n_steps):
        w.plan_now()
        rec["win"].append(w.seek_window())
        rec["state_in"].append(w.state()["state"].astype(np.uint8))
        w.step(1)
        s = w.state()
        for k in ("r", "vel", "inertia", "angle", "angle_vel"):
            rec[k].append(s[k])
        rec["state"].append(s["state"].astype(np.uint8))
        rec["tri"].append(w.find_triangles(s["r"]))
    return {k: np.stack(v) for k, v in rec.items()}


def world_small(seed=7, n_buildings=60, n_agents=2000):
    """512² cells, buildings, mixed unit sizes (UnitTypes.txt: 3/1 and 5/50), 70 % MOVING in 4 groups"""
    rng = np.random.default_rng(seed)
    w = refh.RefWorld(512, 512)
    blds, occ = place_buildings(w, n_buildings, rng, 30, 260, (512, 512))
    w.finalize_map()
    assert w.consistent()
    pos = spawn(rng, n_agents, 150, 900, occ)
    big = rng.random(n_agents) < 0.2
    radius = np.where(big, 5.0, 3.0).astype(np.float32)
    mass = np.where(big, 50.0, 1.0).astype(np.float32)
    player = rng.integers(0, 2, n_agents).astype(np.uint8)
    angle = rng.uniform(-180, 180, n_agents).astype(np.float32)
    for i in range(n_agents):
        w.add_agent(float(pos[i, 0]), float(pos[i, 1]), float(angle[i]), float(radius[i]), float(mass[i]), int(player[i]), 1)
    moving = np.nonzero(rng.random(n_agents) < 0.7)[0].astype(np.int32)
    targets = [(1300.0, 1250.0), (1200.0, 200.0), (100.0, 1300.0), (1400.0, 700.0)]   # all outside the crowd: no arrivals during the recording
    for g, t in enumerate(targets):
        w.issue_paths(moving[g::4], *t)
    w.plan_now()
    return w, dict(radius=radius, mass=mass, player=player, buildings=blds)


def free_point_near(occ, x, y, rng):
    for _ in range(1000):
        if not occ[int(x // 5), int(y // 5)]:
            return float(x), float(y)
        x, y = x + rng.uniform(-40, 40), y + rng.uniform(-40, 40)
    raise RuntimeError("no free point")


def make_map_c3(cells=1024, n_buildings=200, n_groups=64, sub=8, seed=7, name="map_c3.npz"):
    """BASELINE configs 2/3: 1024² cells, 200 octagonal buildings, 64 command groups (8x8 blocks of the map, one random
    target per group).  As PathFinder2::issuePaths2 plans one funnel path per start triangle of a group
    (PathFinder2.cpp:541-606), every group gets one path per sub-block (sub x sub per block) from the sub-block centre
    to the group's target, planned by the reference's PathFinder2: n_groups * sub² paths."""
    rng = np.random.default_rng(seed)
    w = refh.RefWorld(cells, cells)
    blds, occ = place_buildings(w, n_buildings, rng, 12, cells - 12, (cells, cells))
    w.finalize_map()
    assert w.consistent()
    tables = w.tables()
    W = cells * 5.0
    side = int(round(np.sqrt(n_groups)))
    bw = W / side
    sw = bw / sub
    off, wps, pos, ends, starts, group_of = [0], [], [], [], [], []
    targets = []
    for g in range(n_groups):
        bx, by = g % side, g // side
        cx, cy = (bx + 0.5) * bw, (by + 0.5) * bw
        while True:
            ex, ey = free_point_near(occ, rng.uniform(0.1 * W, 0.9 * W), rng.uniform(0.1 * W, 0.9 * W), rng)
            if np.hypot(ex - cx, ey - cy) > 0.3 * W:
                break
        targets.append((ex, ey))
    fails = 0
    for g in range(n_groups):
        bx, by = g % side, g // side
        ex, ey = targets[g]
        for sj in range(sub):
            for si in range(sub):
                sx0, sy0 = bx * bw + (si + 0.5) * sw, by * bw + (sj + 0.5) * sw
                path = None
                for attempt in range(8):
                    sx, sy = free_point_near(occ, sx0 + (rng.uniform(-10, 10) if attempt else 0.0), sy0 + (rng.uniform(-10, 10) if attempt else 0.0), rng)
                    try:
                        path, portals = w.plan_path(sx, sy, ex, ey, 3.0)
                        break
                    except RuntimeError:   # the reference's funnel occasionally runs away (bad_alloc under the rlimit set in main)
                        fails += 1
                if path is None:            # give this sub-block a straight leg to the target
                    path = np.array([[sx, sy], [ex, ey]], np.float32)
                    portals = np.zeros((2, 5), np.float32)
                if len(path) >= 3 and (path[-1] == path[-2]).all():   # setPathOfAgent drops a doubled end, PathFinder2.cpp:333-336
                    path, portals = path[:-1], portals[:-1]
                assert len(path) >= 2 and np.isfinite(path).all() and np.isfinite(portals).all(), (g, path)
                wps.append(path)
                pos.append(portals)
                ends.append([ex, ey])
                starts.append([sx, sy])
                group_of.append(g)
                off.append(off[-1] + len(path))
    np.savez_compressed(os.path.join(OUT, name), **tables, buildings=blds,
                        path_offsets=np.array(off, np.uint32), path_waypoints=np.concatenate(wps).astype(np.float32),
                        path_portals=np.concatenate(pos).astype(np.float32), path_end=np.array(ends, np.float32),
                        path_start=np.array(starts, np.float32), path_group=np.array(group_of, np.int32),
                        path_grid=np.array([side, sub], np.int32))
    print(name, "triangles", len(tables["tri_verts"]), "paths", len(ends), "waypoints", off[-1], "max/path", max(np.diff(off)),
          "planner failures", fails)
    w.close()


def make_map_c5(cells=2048, n_corridors=64, width=6, seg=60, gap=4, x0=64, name="map_c5.npz"):
    """BASELINE config 5: 2048² cells, 64 parallel corridors `width` cells wide running in x, bounded by walls that are
    chains of seg x thick-cell octagonal blocks (corner 1) inserted through the reference's own CDT / Edges code
    (refh_add_building = MapGrid.cpp:1095-1312), `gap` cells apart (alleys between neighbouring corridors). Open plazas at
    both ends. Paths are straight legs (mouth -> far mouth -> plaza), one per corridor and direction."""
    w = refh.RefWorld(cells, cells)
    pitch = cells // n_corridors                 # 32 cells: a corridor and the wall above it
    thick = pitch - width                        # wall blocks fill the space between two corridors
    n_seg = (cells - 2 * x0) // (seg + gap)
    blds = []
    # wall row k lies below corridor k (rows k*pitch + width .. (k+1)*pitch); one more wall row closes the first corridor from above
    for k in range(-1, n_corridors):
        y_lo = k * pitch + width + (pitch - width) // 2 - thick // 2 if k >= 0 else None
        if k < 0:
            continue                             # corridor 0 is bounded above by the map edge region (margin rows)
        cy = k * pitch + width + thick // 2
        if cy + thick // 2 + 2 >= cells:
            continue
        for s_ in range(n_seg):
            cx = x0 + s_ * (seg + gap) + seg // 2
            if w.add_building(cx, cy, seg, thick, 1) != 0:
                raise RuntimeError(w.L.refh_last_error(w.h).decode())
            blds.append((cx, cy, seg, thick))
    w.finalize_map()
    assert w.consistent()
    tables = w.tables()
    W = cells * 5.0
    xa, xb = x0 * 5.0, (x0 + n_seg * (seg + gap) - gap) * 5.0        # corridor mouths
    off, wps, ends, group_of = [0], [], [], []
    for k in range(n_corridors):
        yc = (k * pitch + width / 2.0) * 5.0
        for direction in (0, 1):                 # 0: left -> right, 1: right -> left
            a, b = (xa, xb) if direction == 0 else (xb, xa)
            far = (W - 100.0) if direction == 0 else 100.0
            path = np.array([[a - (40.0 if direction == 0 else -40.0), yc], [a, yc], [b, yc], [far, yc]], np.float32)
            wps.append(path)
            ends.append(path[-1])
            group_of.append(2 * k + direction)
            off.append(off[-1] + len(path))
    nw = off[-1]
    np.savez_compressed(os.path.join(OUT, name), **tables, buildings=np.array(blds, np.int32),
                        path_offsets=np.array(off, np.uint32), path_waypoints=np.concatenate(wps).astype(np.float32),
                        path_portals=np.zeros((nw, 5), np.float32), path_end=np.array(ends, np.float32),
                        path_group=np.array(group_of, np.int32),
                        corridors=np.array([n_corridors, pitch, width, x0, n_seg, seg, gap], np.int32))
    print(name, "triangles", len(tables["tri_verts"]), "wall blocks", len(blds), "paths", len(ends))
    w.close()


def make_map_insert(name="map_insert.npz", seed=11):
    """f3: the table set before and after ONE building insert (Game::updateTriangulation, Game.cpp:38-48): 512² cells, 40
    buildings, then a 41st inserted into the live CDT through the same reference code; the reference's own updateCellGrid
    ran after both. tests/test_gpu_parity.py feeds the difference to rts_update_map_region."""
    rng = np.random.default_rng(seed)
    w = refh.RefWorld(512, 512)
    blds, occ = place_buildings(w, 40, rng, 30, 260, (512, 512))
    w.finalize_map()
    before = w.tables()
    lf_before = w.last_found()
    while True:
        cx, cy = (int(v) for v in rng.integers(60, 230, 2))
        if not occ[cx - 7:cx + 7, cy - 7:cy + 7].any():
            break
    assert w.add_building(cx, cy, 8, 8, 1) == 0
    w.finalize_map()
    assert w.consistent()
    after = w.tables()
    np.savez_compressed(os.path.join(OUT, name), **{"a_" + k: v for k, v in before.items()}, **{"b_" + k: v for k, v in after.items()},
                        new_building=np.array([cx, cy, 8, 8], np.int32), last_found_before=np.array(lf_before))
    print(name, "triangles", len(before["tri_verts"]), "->", len(after["tri_verts"]), "new building at", cx, cy)
    w.close()


def main():
    import resource
    resource.setrlimit(resource.RLIMIT_AS, (12 << 30, 12 << 30))   # a runaway reference planner must not take the box down
    assert refh.available(), "build oracle/_ref first: make -C oracle ref"
    if len(sys.argv) > 1 and sys.argv[1] == "c5":                  # only the corridor map (the other fixtures are unchanged)
        make_map_c5()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "insert":              # only the before / after pair of a building insert
        make_map_insert()
        return
    L = refh.load()

    # ---- known answers of the reference's own tests ------------------------------------------------
    xy = np.array([[25, 25], [40, 40], [24, 24]], np.float32)
    rad = np.full(3, 5, np.float32)
    first = C.c_int(-1)
    kat = {}
    kat["ns_count_agent0"] = L.refh_seek_ns_test(100, 10, xy, rad, 3, 0, C.byref(first))   # test_neighbour_search.cc:36
    kat["ns_first_agent0"] = first.value                                                   # :37
    kat["ns_count_agent1"] = L.refh_seek_ns_test(100, 10, xy, rad, 3, 1, C.byref(first))   # :39
    kat["ns_full_r10"] = L.refh_seek_ns_full(100, 10, xy, rad, 3, 39, 39, 10)              # :63
    kat["ns_full_r1"] = L.refh_seek_ns_full(100, 10, xy, rad, 3, 39, 39, 1)                # :65
    w = refh.RefWorld(512, 512)
    kat["ntri_boundary"] = w.L.refh_n_triangles(w.h)
    w.L.refh_insert_vertex(w.h, 50, 50, 0)
    kat["ntri_after_vertex"] = w.L.refh_n_triangles(w.h)                                   # test_triangulation.cc:28
    kat["find_54_54"] = int(w.L.refh_find_triangle_int(w.h, 54, 54))                       # test_triangulation.cc:114
    tb = w.tables()   # cell2tri is all -1 here (updateCellGrid not called, as in the test)
    assert kat == {"ns_count_agent0": 1, "ns_first_agent0": 2, "ns_count_agent1": 0, "ns_full_r10": 1, "ns_full_r1": 0,
                   "ntri_boundary": 9, "ntri_after_vertex": 11, "find_54_54": 4}, kat
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **{k: np.array(v) for k, v in kat.items()},
                        **{"tf_" + k: v for k, v in tb.items()})
    w.close()

    # ---- multi-step frame parity on a 512² map with buildings -------------------------------------------
    w, info = world_small()
    tables = w.tables()
    w.step(60)                       # warm up: inertia, overlaps, slid waypoints
    s0 = w.state()
    rec = record_steps(w, 6)
    np.savez_compressed(os.path.join(OUT, "frame_small.npz"), **tables, **info,
                        r0=s0["r"], inertia0=s0["inertia"], angle0=s0["angle"], angle_vel0=s0["angle_vel"],
                        **{"x_" + k: v for k, v in rec.items()})
    # point queries on the same map
    rng = np.random.default_rng(11)
    lattice = np.stack(np.meshgrid(np.arange(0, 1400, 7), np.arange(0, 1400, 7)), -1).reshape(-1, 2).astype(np.float32)
    pts = np.concatenate([lattice, rng.uniform(0, 2560, (20000, 2)).astype(np.float32)])
    tri = w.find_triangles(pts)
    # wall contacts: points around every building, radius 3 and 5
    bl = info["buildings"]
    q = []
    for cx, cy, n, m in bl:
        q.append(rng.uniform([5 * (cx - n / 2) - 8, 5 * (cy - m / 2) - 8], [5 * (cx + n / 2) + 8, 5 * (cy + m / 2) + 8], (60, 2)))
    q = np.concatenate(q).astype(np.float32)
    qr = np.where(rng.random(len(q)) < 0.5, 3.0, 5.0).astype(np.float32)
    cap = 8
    cnt = np.zeros(len(q), np.uint32)
    ce = np.zeros((len(q), cap, 5), np.float32)
    for i in range(len(q)):
        e = w.contact_edges(float(q[i, 0]), float(q[i, 1]), float(qr[i]), cap)
        cnt[i] = len(e)
        ce[i, :len(e)] = e
    # Grid::coordToCell on the three grids, incl. cell-boundary points
    cpts = np.concatenate([rng.uniform(0, 2560, (4000, 2)), np.arange(0, 2560, 5.0)[:, None].repeat(2, 1),
                           np.nextafter(np.arange(60, 2560, 60.0, dtype=np.float32), np.float32(0))[:, None].repeat(2, 1),
                           np.arange(0, 2560, 30.0)[:, None] + np.array([[0.0, 17.3]])]).astype(np.float32)
    cells = {f"cell_{nx}_{int(cs)}": refh.coord_to_cell(nx, nx, cs, cpts) for nx, cs in ((43, 60.0), (86, 30.0), (512, 5.0), (64, 40.0))}
    np.savez_compressed(os.path.join(OUT, "queries_small.npz"), tri_pts=pts, tri_expected=tri, wall_pts=q, wall_radius=qr,
                        wall_counts=cnt, wall_edges=ce, cell_pts=cpts, **cells)
    w.close()

    # ---- dense crowd: C3 density (0.038 agents / unit²), overlapping, one wall-free frame + one with walls --------
    rng = np.random.default_rng(1236)
    w = refh.RefWorld(512, 512)
    blds, occ = place_buildings(w, 6, rng, 45, 100, (512, 512))
    w.finalize_map()
    n = 5000
    pos = spawn(rng, n, 200.0, 562.0, occ)
    big = rng.random(n) < 0.2
    radius = np.where(big, 5.0, 3.0).astype(np.float32)
    mass = np.where(big, 50.0, 1.0).astype(np.float32)
    for i in range(n):
        w.add_agent(float(pos[i, 0]), float(pos[i, 1]), 0.0, float(radius[i]), float(mass[i]), 0, 1)
    ids = np.arange(n, dtype=np.int32)
    w.issue_paths(ids[: n // 2], 1500.0, 1400.0)
    w.plan_now()
    tables = w.tables()
    s0 = w.state()
    rec = record_steps(w, 3)
    np.savez_compressed(os.path.join(OUT, "frame_dense.npz"), **tables, radius=radius, mass=mass,
                        player=np.zeros(n, np.uint8), buildings=blds,
                        r0=s0["r"], inertia0=s0["inertia"], angle0=s0["angle"], angle_vel0=s0["angle_vel"],
                        **{"x_" + k: v for k, v in rec.items()})
    w.close()
    # ---- empty 512² map (BASELINE config 1) -----------------------------------------------------------
    w = refh.RefWorld(512, 512)
    w.finalize_map()
    np.savez_compressed(os.path.join(OUT, "map_empty.npz"), **w.tables(), buildings=np.zeros((0, 4), np.int32))
    w.close()
    if "--with-c3" in sys.argv:      # ~8 minutes: 4096 planner queries, a few of which run away until the rlimit stops them
        make_map_c3()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
