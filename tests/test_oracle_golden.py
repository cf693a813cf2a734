"""CPU: the plain-C oracle (oracle/rts_oracle.c) against fixtures produced by the verbatim reference
(tests/golden/make_golden.py).  These pin the oracle; the GPU tests then compare against the oracle."""
import numpy as np
import pytest

from oracle import port
from tests import helpers as H


def test_known_answers_of_reference_tests():
    k = H.load("kat.npz")
    # src/Tests/test_neighbour_search.cc:36-39,63-65 and src/Tests/test_triangulation.cc:28,114
    assert (int(k["ns_count_agent0"]), int(k["ns_first_agent0"]), int(k["ns_count_agent1"])) == (1, 2, 0)
    assert (int(k["ns_full_r10"]), int(k["ns_full_r1"])) == (1, 0)
    assert int(k["ntri_after_vertex"]) == 11
    # TestTriangleFinder: boundary 2560² + vertex (50,50), cell cache empty -> findTriangle({54,54}) == 4
    P = port.default_params(512, 512)
    ow = port.OracleWorld(P, H.tables_of(k, "tf_"))
    assert int(ow.find_triangles(np.array([[54.0, 54.0]], np.float32))[0]) == int(k["find_54_54"]) == 4


def test_seek_neighbour_search_known_answer():
    """TestSmallestDistance (test_neighbour_search.cc:8-40): 100² box, cell 10, cutoff 10*(ri+rj)^2.
    The physics strategy shares the cell walk; restated here with r_max2 = 10*(5+5)^2 on a 10-unit grid."""
    P = port.default_params(20, 20)
    P.ns_cell_size = 10.0
    P.ns_nx = P.ns_ny = 11
    P.r_max2 = 10.0 * (5 + 5) ** 2
    ow = port.OracleWorld(P, H.tables_of(H.load("kat.npz"), "tf_"))
    r = np.array([[25, 25], [40, 40], [24, 24]], np.float32)
    # cutoff 1000 > (15*sqrt2)^2=450 would also catch agent 1 if cells were adjacent: they are not (cells 2 and 4)
    assert list(ow.neighbours(r, 0)) == [2]
    assert list(ow.neighbours(r, 1)) == []


def test_point_queries():
    q = H.load("queries_small.npz")
    fx = H.load("frame_small.npz")
    P = port.default_params(512, 512)
    ow = port.OracleWorld(P, H.tables_of(fx))
    H.assert_same_bits(ow.find_triangles(q["tri_pts"]), q["tri_expected"], "findTriangle")
    cnt, edges = ow.contact_edges(q["wall_pts"], q["wall_radius"], cap=q["wall_edges"].shape[1])
    H.assert_same_bits(cnt, q["wall_counts"], "calcContactEdgesO count")
    H.assert_same_bits(edges, q["wall_edges"], "calcContactEdgesO edges")
    assert (q["wall_counts"] > 0).sum() > 100 and (q["wall_counts"] > 1).sum() > 5
    for nx, cs in ((43, 60.0), (86, 30.0), (512, 5.0), (64, 40.0)):
        H.assert_same_bits(port.coord_to_cell(q["cell_pts"], nx, nx, cs), q[f"cell_{nx}_{int(cs)}"].astype(np.uint32),
                           f"coordToCell {nx}x{cs}")


@pytest.mark.parametrize("name", ["frame_small.npz", "frame_dense.npz"])
def test_frames_bit_exact(name):
    fx = H.load(name)
    P = port.default_params(512, 512)
    tables = H.tables_of(fx)
    prev = None
    n_frames = len(fx["x_r"])
    slid = advanced = 0
    for s in range(n_frames):
        a, paths = H.frame_agents(fx, s, prev)
        ow = port.OracleWorld(P, tables, paths)
        ow.step(a)
        for k in ("r", "vel", "inertia", "angle", "angle_vel"):
            H.assert_same_bits(a[k], fx["x_" + k][s], f"{name} frame {s} {k}")
        H.assert_same_bits(a["tri_idx"], fx["x_tri"][s], f"{name} frame {s} tri")
        H.assert_same_bits(a["state"], fx["x_state"][s], f"{name} frame {s} state")
        slid += int((a["r_next"] != fx["x_win"][s][:, 0:2]).any(1).sum())
        advanced += int((a["waypoint"] > 0).sum())
        prev = a
    assert np.isfinite(prev["r"]).all()
    if name == "frame_small.npz":
        assert slid + advanced > 0      # the fixture exercises portal slides / waypoint shifts


def test_ghosts_are_neighbours_only():
    fx = H.load("frame_dense.npz")
    P = port.default_params(512, 512)
    a, paths = H.frame_agents(fx, 0)
    ow = port.OracleWorld(P, H.tables_of(fx), paths)
    ref = {k: v.copy() for k, v in a.items()}
    ow.step(ref)
    ghost = (np.arange(len(a["r"])) % 3 == 0).astype(np.uint8)
    b = {k: v.copy() for k, v in a.items()}
    ow.step(b, ghost=ghost)
    own = ghost == 0
    H.assert_same_bits(b["r"][own], ref["r"][own], "owned agents unaffected by ghost marking")
    H.assert_same_bits(b["r"][~own], a["r"][~own], "ghosts are not moved")


@pytest.mark.parametrize("name,cells,last_found", [("frame_small.npz", (512, 512), 0), ("frame_dense.npz", (512, 512), 0),
                                                   ("map_empty.npz", (512, 512), 2), ("map_c3.npz", (1024, 1024), 0),
                                                   ("map_c5.npz", (2048, 2048), 0)])
def test_cell_grid_restatement_reproduces_recorded_tables(name, cells, last_found):
    """Triangulation::updateCellGrid (Triangulation.cpp:1829-1864): the tables in the fixtures were written by the
    reference itself; 95 / 2 / 64 / 196 (and, on the corridor map, many more) of their cell centres lie on triangle edges or vertices, where the entry depends
    on the zig-zag walk (map_empty: the reference's last_found_tri_ind_ was 2 when it built the table)."""
    tables = H.tables_of(H.load(name))
    ow = port.OracleWorld(port.default_params(*cells), tables)
    assert np.array_equal(ow.build_cell_grid(last_found), np.asarray(tables["cell2tri"], np.uint32))
