"""CPU: host-side logic that needs no device — ABI struct helpers, scenario generators, map tiling."""
import numpy as np
import pytest

from oracle import port
from rtsengine_b200 import abi, scenario
from tests import helpers as H


def test_agent_view_checks_shapes_and_dtypes():
    keep = []
    v = abi.agent_view({"r": np.zeros((5, 2), np.float32), "state": np.ones(5, np.uint8)}, 5, keep)
    assert v.r and v.state and not v.inertia
    with pytest.raises(ValueError):
        abi.agent_view({"r": np.zeros((4, 2), np.float32)}, 5, [])
    t = abi.triangles_array(np.arange(12).reshape(2, 6), np.array([[1, 2, 0xFFFFFFFF], [0, 0xFFFFFFFF, 0xFFFFFFFF]]))
    assert t["vx"].tolist() == [[0, 2, 4], [6, 8, 10]] and t["vy"].tolist() == [[1, 3, 5], [7, 9, 11]]
    with pytest.raises(ValueError):
        abi.path_table({"offsets": np.array([0, 2], np.uint32), "waypoints": np.zeros((1, 2)), "portals": np.zeros((2, 5)),
                        "path_end": np.zeros((1, 2))}, [])
    pt = abi.path_table({"offsets": np.array([0, 1, 3], np.uint32), "waypoints": np.zeros((3, 2)), "portals": np.zeros((3, 5)),
                         "path_end": np.zeros((2, 2)), "group": np.array([4, 4])}, [])
    assert (pt.n_paths, pt.n_groups) == (2, 5)


def test_config3_population_and_paths():
    fx = scenario.load_map("map_c3.npz")
    a, paths = scenario.config3_agents(fx, n=50_000, seed=3)
    assert len(paths["path_end"]) == 4096 and paths["offsets"][-1] == len(paths["waypoints"])
    assert a["path_id"].min() >= 0 and a["path_id"].max() < 4096 and (a["state"] == abi.MOVING).all()
    # nobody spawns inside a building (one-cell margin), everybody inside the map
    occ = scenario.occupancy(fx, 0)
    assert not occ[(a["r"][:, 0] // 5).astype(int), (a["r"][:, 1] // 5).astype(int)].any()
    assert (a["r"] > 0).all() and (a["r"] < 5120).all()
    # an agent follows the path planned from its own start region, towards its group's destination
    pid = a["path_id"]
    start = fx["path_start"][pid]
    assert np.linalg.norm(start - a["r"], axis=1).max() < 80 * np.sqrt(2) + 60
    side, sub = (int(v) for v in fx["path_grid"])
    assert (fx["path_group"][pid] == pid // (sub * sub)).all()
    # every path ends at its destination and has finite portals
    last = paths["offsets"][1:] - 1
    assert np.allclose(paths["waypoints"][last], paths["path_end"])
    assert np.isfinite(paths["portals"]).all()


def test_config1_and_config2_populations():
    """BASELINE configs[0] (2 000 agents in [200,1200]^2 of the empty 512^2 map, one group bound for (2000,2000)) and
    configs[1] (100k agents, the unit mix of UnitTypes.txt, half of them MOVING in 16 command groups, right up to the walls)"""
    a, paths = scenario.config1_agents()
    assert len(a["r"]) == 2000 and (a["r"] >= 200).all() and (a["r"] <= 1200).all()
    assert (a["state"] == abi.MOVING).all() and (a["path_id"] == 0).all() and (a["radius"] == 3).all() and (a["mass"] == 1).all()
    assert paths["path_end"].tolist() == [[2000.0, 2000.0]] and paths["offsets"].tolist() == [0, 1]
    assert len(np.unique(a["r"], axis=0)) == 2000                                    # no two agents coincide
    fx = scenario.load_map("map_c3.npz")
    a, paths = scenario.config2_agents(fx, n=100_000)
    big = a["unit_type"] == 1
    assert 0.19 < big.mean() < 0.21
    assert (a["radius"][big] == 5).all() and (a["mass"][big] == 50).all() and (a["radius"][~big] == 3).all() and (a["mass"][~big] == 1).all()
    moving = a["state"] == abi.MOVING
    assert 0.49 < moving.mean() < 0.51 and (a["path_id"][~moving] == -1).all() and (a["path_id"][moving] >= 0).all()
    assert len(np.unique(fx["path_group"][a["path_id"][moving]])) == 16
    occ = scenario.occupancy(fx, 0)
    assert not occ[(a["r"][:, 0] // 5).astype(int), (a["r"][:, 1] // 5).astype(int)].any()
    # ... and some of them start in contact with a wall (the oracle's wall query as the checker)
    ow = port.OracleWorld(port.default_params(1024, 1024), scenario.tables_of(fx))
    cnt, _ = ow.contact_edges(a["r"][:20000], a["radius"][:20000])
    assert 20 < (cnt > 0).sum() < 2000


def test_config5_population_and_paths():
    """BASELINE configs[4]: the corridor population — free space only, everyone MOVING, the two halves of the map heading
    through the corridor nearest to them in opposite directions (counter-flow), each agent already on the leg of its path
    that lies ahead of it"""
    fx = scenario.load_map("map_c5.npz")
    n_corr, pitch, width, x0, n_seg, seg, gap = (int(v) for v in fx["corridors"])
    a, paths = scenario.config5_agents(fx, n=60_000, seed=5)
    W = int(fx["n_cells"][0]) * 5.0
    assert len(paths["path_end"]) == 2 * n_corr and paths["offsets"][-1] == len(paths["waypoints"]) == 4 * 2 * n_corr
    assert (a["state"] == abi.MOVING).all() and (a["radius"] == 3).all() and (a["mass"] == 1).all()
    occ = scenario.occupancy(fx, 0)
    x, y = a["r"][:, 0], a["r"][:, 1]
    assert not occ[(x // 5).astype(int), (y // 5).astype(int)].any()
    assert (a["r"] > 0).all() and (a["r"] < W).all()
    pid, wp = a["path_id"], a["waypoint"]
    assert pid.min() == 0 and pid.max() == 2 * n_corr - 1
    heads_left = (pid % 2).astype(bool)
    assert (heads_left == (x >= W / 2)).all() and 0.4 < heads_left.mean() < 0.6
    # the corridor of an agent's path is the one nearest to it, and both directions share every corridor
    lane_y = paths["waypoints"][paths["offsets"][pid] + 1][:, 1]
    lanes = np.unique(paths["waypoints"][paths["offsets"][:-1] + 1][:, 1])
    assert len(lanes) == n_corr
    nearest = np.abs(y[:, None] - lanes[None, :]).min(axis=1)
    assert np.abs(np.abs(lane_y - y) - nearest).max() < 1e-3
    assert len(np.unique(pid[~heads_left] // 2)) == n_corr and len(np.unique(pid[heads_left] // 2)) == n_corr
    # the waypoint an agent aims at lies ahead of it in its direction of travel
    target = paths["waypoints"][paths["offsets"][pid] + wp]
    ahead = np.where(heads_left, target[:, 0] <= x + 1e-3, target[:, 0] >= x - 1e-3)
    assert wp.min() >= 1 and wp.max() <= 3 and ahead.mean() > 0.999
    last = paths["offsets"][1:] - 1
    assert np.array_equal(paths["waypoints"][last], paths["path_end"]) and np.isfinite(paths["portals"]).all()


def test_tiled_map_is_the_tile_map_shifted():
    """a 2x3 tiling of the config-3 map must answer point queries like the single tile (oracle as the checker)"""
    tile = scenario.load_map("map_c3.npz")
    big = scenario.tile_map(tile, 2, 3)
    assert big["n_cells"].tolist() == [2048, 3072]
    assert [len(big[f"line_off{o}"]) - 1 for o in range(4)] == [3073, 5121, 2049, 5121]      # Edges.cpp:6-19
    ow_big = port.OracleWorld(port.default_params(2048, 3072), scenario.tables_of(big))
    ow_one = port.OracleWorld(port.default_params(1024, 1024), scenario.tables_of(tile))
    rng = np.random.default_rng(0)
    local = rng.uniform(5, 5115, (20000, 2)).astype(np.float32)
    bl = tile["buildings"]
    near = np.concatenate([rng.uniform([5 * (cx - n / 2) - 6, 5 * (cy - m / 2) - 6], [5 * (cx + n / 2) + 6, 5 * (cy + m / 2) + 6], (20, 2))
                           for cx, cy, n, m in bl]).astype(np.float32)
    rad = np.where(rng.random(len(near)) < 0.5, 3.0, 5.0).astype(np.float32)
    c1, e1 = ow_one.contact_edges(near, rad)
    assert (c1 > 0).sum() > 200
    for t, (i, j) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]):
        off = np.array([5120.0 * i, 5120.0 * j], np.float32)
        H.assert_same_bits(ow_big.find_triangles(local + off), big["tile_tri_index"][t][ow_one.find_triangles(local)], f"tile {t} triangles")
        cb, eb = ow_big.contact_edges(near + off, rad)
        H.assert_same_bits(cb, c1, f"tile {t} wall contact counts")
        H.assert_same_bits(eb[..., 2:], e1[..., 2:], f"tile {t} wall tangents / lengths")


def test_reference_arm_of_the_bench_emits_the_contract_line():
    """`bench.py --impl reference` (CPU only): one JSON line with the keys the driver reads; the verbatim reference when
    oracle/_ref exists, the oracle port otherwise."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports to its workers
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "agent-updates/sec" and j["unit"] == "agent-updates/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 2
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["value"] == j["value"]
    assert j["cpu_baseline"]["cores"] == (os.cpu_count() or 1)       # every host thread, whatever OMP_NUM_THREADS said
    if j["cpu_baseline"]["kind"] == "reference" and (os.cpu_count() or 1) >= 6:
        stock = j["cpu_baseline"]["stock_threads"]                   # the reference's hard-coded 6 threads beside it
        assert stock["cores"] == 6 and stock["value"] > 0
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the record says what really ran: the reference cannot hold config 3 (5 000-agent cap), so the arm runs its density
    assert j["config"]["same_config_as_gpu_arm"] is False and j["config"]["agents"] in (5000, 200000)
    assert "config-3" in j["config"]["workload"] or "config 3" in j["config"]["workload"]


def test_rebalanced_bounds_follow_the_histogram_within_the_migrant_capacity():
    """slab rebalancing (host side): boundaries move toward the agent-count quantiles, at most max_shift rows, never more
    rows than whose agents fit the migrant buffers, and not at all inside the tolerance"""
    import numpy as np
    from rtsengine_b200 import slabs
    hist = np.array([100] * 10 + [400] * 10 + [100] * 10)
    assert slabs.rebalanced_bounds(hist, [0, 10, 20, 30], cap_migrants=2000) == [0, 13, 18, 30]
    assert slabs.rebalanced_bounds(hist, [0, 10, 20, 30], cap_migrants=500) == [0, 11, 19, 30]     # one 400-agent row fits, two do not
    assert slabs.rebalanced_bounds(np.array([100] * 30), [0, 10, 20, 30], cap_migrants=500) == [0, 10, 20, 30]
    b = slabs.rebalanced_bounds(np.array([0] * 29 + [1000]), [0, 10, 20, 30], cap_migrants=10 ** 6, max_shift=100)
    assert b[0] == 0 and b[-1] == 30 and all(b[i] < b[i + 1] for i in range(3))
