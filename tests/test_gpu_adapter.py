"""GPU: the C++ drop-in adapter (rtsengine_b200/adapter/GpuAgentSystem.h) inside the reference's own frame loop.

oracle/_ref/adapter_demo (built here by `make -C oracle demo`, it links the verbatim reference translation units) runs
the reference's PhysicsSystem + SeekSystem and the GPU slots side by side — same map, same units, same PathFinder2
planning, same ECSystem::update orchestration — and compares the ECS blackboards bit for bit after every frame."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "adapter_demo")


@pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/adapter_demo not built (needs /root/reference at build time)")
@pytest.mark.parametrize("agents,buildings,frames,seed", [(2000, 40, 150, 7), (3000, 60, 80, 11), (4800, 50, 40, 2)])   # (the last one ends before the insert)
def test_system2_adapter_matches_reference_systems(agents, buildings, frames, seed):
    # (the demo hands the reference planner its pending groups 8 at a time: PathFinder2 indexes 12-entry per-thread
    # arrays with worker ids up to hardware_concurrency()-2, PathFinder2.h:161-162 / PathFinder2.cpp:90)
    cmd = [DEMO, str(agents), str(buildings), str(frames), str(seed)]
    # A hang is a failure, never a skip. (The reference's PathFinder2 does not terminate on some maps with many buildings;
    # the three cases here are maps it handles — `RTS_DEMO_REFERENCE_ONLY=1 <cmd>` runs world A alone to tell the two apart.)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert line, out.stdout + out.stderr
    res = json.loads(line[-1])
    assert "error" not in res, res
    assert res["cell_grid_mismatches"] == 0, res          # rts_build_cell_grid == Triangulation::updateCellGrid
    assert res["compared"] == agents * frames
    assert res["mismatches"] == 0, res
    assert res["moving_at_end"] > 0.5 * agents          # the move commands were still being followed
    # f3: the building inserted at frame 40 reached the GPU slot as a difference (rts_update_map_region), the patched
    # cell -> triangle cache equals the reference's full rebuild (counted in cell_grid_mismatches) and the frames after it match
    if frames > 40:
        mu = res["map_update"]
        assert mu["changed_wall_lines"] == 8 and 8 <= mu["changed_triangles"] <= 200, mu
        assert 0 < mu["bytes"] < 0.10 * mu["full_upload_bytes"], mu
    # dirty tracking: after the first frame only the entities of the late command (frame 25) were re-sent, not the blackboard
    if frames > 25:
        assert res["late_command"] > 0 and res["entities_resent_after_frame0"] == res["late_command"], res
    else:
        assert res["entities_resent_after_frame0"] == 0, res
    assert out.returncode == 0
