// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Drop-in demonstration: the SAME frame loop (ECSystem::update, src/ECS.cpp:65-113, restricted to the PHYSICS and
// PATHFINDING slots) is driven twice, side by side, from identical initial states on identical maps:
//
//   world A: the reference's own PhysicsSystem + SeekSystem (verbatim translation units, CPU)
//   world B: GpuPhysicsSlot + GpuSeekSlot of rtsengine_b200/adapter/GpuAgentSystem.h (librts_b200.so, GPU),
//            installed in the same System2 slots, with the reference's own PathFinder2 planning for both
//
// and the blackboards (std::array<SharedData, N_MAX_ENTITIES>) are compared bit for bit after every frame.
// Built by `make -C oracle demo` into oracle/_ref/adapter_demo; run by tests/test_gpu_adapter.py on the GPU box.
//
// usage: adapter_demo <n_agents> <n_buildings> <frames> [seed]      prints one JSON line

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <vector>
#include <omp.h>
// every standard header the reference pulls in, before the access-specifier trick below
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <map>
#include <numbers>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <typeindex>
#include <typeinfo>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>

#define private public
#define protected public
#include "Systems/PhysicsSystem.h"
#include "Systems/SeekSystem.h"
#include "PathFinding/PathFinder2.h"
#include "PathFinding/Triangulation.h"
#undef private
#undef protected
#include "Utils/Grid.h"
#include "Edges.h"
#include "GpuAgentSystem.h"

extern "C" {
void* refh_create(int, int);
int refh_add_building(void*, int, int, int, int, int);
int refh_finalize_map(void*);
}

// the harness' world struct (oracle/ref_harness.cpp) — same layout, declared here to reach its members
struct RefWorld {
    sf::Vector2i n_cells;
    std::unique_ptr<SearchGrid> tri_grid;
    std::unique_ptr<Triangulation> cdt;
    std::unique_ptr<Edges> edges;
    std::unique_ptr<PhysicsSystem> ps;
    std::unique_ptr<SeekSystem> ss;
    std::unique_ptr<PathFinder2> pf;
    std::unique_ptr<std::array<SharedData, N_MAX_ENTITIES>> shared;
    std::vector<Entity> active;
    std::vector<sf::Vector2f> vel_prezero;
    int n_agents = 0;
    std::string last_error;
};

// see plan_pending in oracle/ref_harness.cpp: the reference's planner must not get more than 12 groups at once
static void plan_pending(PathFinder2& pf, std::vector<PathFinderComponent>& comps, std::array<int, N_MAX_ENTITIES>& e2c) {
    std::remove_reference_t<decltype(pf.to_update_groups_)> all;
    all.swap(pf.to_update_groups_);
    for (size_t i = 0; i < all.size(); i += 8) {
        pf.to_update_groups_.assign(all.begin() + i, all.begin() + std::min(i + 8, all.size()));
        if (getenv("RTS_DEMO_TRACE")) fprintf(stderr, "planning groups %zu.. of %zu\n", i, all.size());
        pf.updatePaths(comps, e2c, 5000);
    }
}

static bool same_bits(float a, float b) {
    uint32_t x, y;
    std::memcpy(&x, &a, 4);
    std::memcpy(&y, &b, 4);
    return x == y || (a == 0.f && b == 0.f) || (a != a && b != b);
}

int main(int argc, char** argv) {
    const int n_agents = argc > 1 ? atoi(argv[1]) : 2000;
    const int n_buildings = argc > 2 ? atoi(argv[2]) : 40;
    const int frames = argc > 3 ? atoi(argv[3]) : 120;
    const unsigned seed = argc > 4 ? atoi(argv[4]) : 7;

    // ---- two identical worlds (map, planner) built by the reference's own code
    std::mt19937 rng(seed);
    std::vector<std::array<int, 4>> blds;
    std::vector<char> occ(512 * 512, 0);
    auto blocked = [&](int x0, int y0, int x1, int y1) {
        for (int x = std::max(x0, 0); x < std::min(x1, 512); ++x)
            for (int y = std::max(y0, 0); y < std::min(y1, 512); ++y)
                if (occ[x * 512 + y]) return true;
        return false;
    };
    while ((int)blds.size() < n_buildings) {
        const int cx = 30 + rng() % 230, cy = 30 + rng() % 230;
        const bool big = rng() % 2;
        const int n = big ? 8 : 4, m = big ? 8 : 6;
        const int x0 = cx - n / 2, y0 = cy - m / 2;
        if (blocked(x0 - 2, y0 - 2, x0 + n + 2, y0 + m + 2)) continue;
        for (int x = x0 - 1; x < x0 + n + 1; ++x)
            for (int y = y0 - 1; y < y0 + m + 1; ++y) occ[x * 512 + y] = 1;
        blds.push_back({cx, cy, n, m});
    }
    RefWorld* W[2];
    for (int k = 0; k < 2; ++k) {
        W[k] = static_cast<RefWorld*>(refh_create(512, 512));
        for (auto& b : blds)
            if (refh_add_building(W[k], b[0], b[1], b[2], b[3], 1) != 0) { printf("{\"error\": \"building\"}\n"); return 1; }
        if (refh_finalize_map(W[k]) != 0) { printf("{\"error\": \"finalize\"}\n"); return 1; }
    }
    RefWorld& A = *W[0];
    RefWorld& B = *W[1];
    long cell_grid_mismatches = 0;

    // ---- world B: the GPU slots in place of the CPU systems (what INTEGRATION.md §2 adds to ECSystem's constructor)
    RtsParams params;
    rts_default_params(&params, 512, 512);
    auto phys = std::make_unique<GpuPhysicsSlot>(ComponentID::PHYSICS);
    phys->p_comps_ = std::make_shared<ComponentArray<PhysicsComponent>>();
    std::unique_ptr<GpuSeekSlot> seek;
    const bool gpu = getenv("RTS_DEMO_REFERENCE_ONLY") == nullptr;   // debugging aid: world A alone, no device needed
    if (gpu) try {
        seek = std::make_unique<GpuSeekSlot>(ComponentID::PATHFINDING, params, 0);
    } catch (std::exception& e) {
        printf("{\"error\": \"%s\"}\n", e.what());
        return 2;
    }
    if (gpu) seek->p_comps_ = std::make_shared<ComponentArray<PathFinderComponent>>();
    if (gpu) seek->physics = phys.get();
    A.cdt->last_found_tri_ind_ = 0;
    A.cdt->updateCellGrid();
    if (gpu) {   // Game::updateTriangulation hook: flatten Edges + CDT and upload
        auto& cdt = *B.cdt;
        std::vector<RtsTriangle> tris(cdt.triangles_.size());
        for (size_t t = 0; t < tris.size(); ++t)
            for (int k = 0; k < 3; ++k) {
                tris[t].vx[k] = cdt.triangles_[t].verts[k].x;
                tris[t].vy[k] = cdt.triangles_[t].verts[k].y;
                tris[t].neighbours[k] = cdt.triangles_[t].neighbours[k];
                tris[t].is_constrained[k] = cdt.triangles_[t].is_constrained[k];
            }
        RtsWallTable walls{};
        std::vector<RtsEdge> e[4];
        std::vector<uint32_t> off[4];
        for (int o = 0; o < 4; ++o) {
            for (auto& line : B.edges->edge_finders_[o]->edge_array2_) {
                off[o].push_back((uint32_t)e[o].size());
                for (auto& s : line) e[o].push_back({s.from.x, s.from.y, s.t.x, s.t.y, s.l});
            }
            off[o].push_back((uint32_t)e[o].size());
            walls.edges[o] = e[o].data();
            walls.line_offsets[o] = off[o].data();
            walls.n_lines[o] = (uint32_t)off[o].size() - 1;
        }
        // the findTriangle cache: the host loop (Triangulation::updateCellGrid, above) in world A, the device builder in world B
        seek->uploadMap(tris.data(), (uint32_t)tris.size(), nullptr, walls);
        std::vector<uint32_t> built(cdt.cell2triangle_ind_.size());
        if (rts_build_cell_grid(seek->h, 0u, built.data()) != RTS_OK) { printf("{\"error\": \"rts_build_cell_grid\"}\n"); return 2; }
        for (size_t c = 0; c < built.size(); ++c) {
            cell_grid_mismatches += built[c] != (uint32_t)A.cdt->cell2triangle_ind_[c];
            cdt.cell2triangle_ind_[c] = (TriInd)built[c];   // world B's host planner reads the device-built table
        }
    }

    // ---- units (UnitInitializer::initializeUnit): identical in both worlds
    std::uniform_real_distribution<float> U(150.f, 900.f), Ang(-180.f, 180.f);
    for (int i = 0; i < n_agents; ++i) {
        sf::Vector2f pos;
        do { pos = {U(rng), U(rng)}; } while (occ[int(pos.x / 5) * 512 + int(pos.y / 5)]);
        const bool big = rng() % 5 == 0;
        TransformComponent tc;
        tc.r = pos;
        tc.angle = Ang(rng);
        PhysicsComponent pc;
        pc.transform = tc;
        pc.radius = big ? 5.f : 3.f;
        pc.mass = big ? 50.f : 1.f;
        pc.player_ind = i % 2;
        PathFinderComponent pfc;
        pfc.transform = tc;
        pfc.radius = pc.radius;
        Entity ent(0, i);
        SharedData sd;
        sd.transform = tc;
        sd.state = MoveState::STANDING;
        A.shared->at(i) = sd;
        B.shared->at(i) = sd;
        A.ps->addComponent(ent, pc);
        A.ss->addComponent(ent, pfc);
        if (gpu) phys->addComponent(ent, pc);
        if (gpu) seek->addComponent(ent, pfc);
        A.active.push_back(ent);
        B.active.push_back(ent);
    }
    if (gpu) seek->markPopulationDirty();

    // ---- move commands (Game.cpp:211-231): 70 % of the units in three groups
    std::vector<std::vector<int>> groups(3);
    for (int i = 0; i < n_agents; ++i)
        if (rng() % 10 < 7) groups[i % 3].push_back(i);
    const sf::Vector2f targets[3] = {{1500.f, 1400.f}, {1450.f, 250.f}, {200.f, 1500.f}};
    for (int g = 0; g < 3; ++g) {
        for (int e : groups[g]) { A.shared->at(e).state = MoveState::MOVING; B.shared->at(e).state = MoveState::MOVING; }
        A.ss->issuePaths2(groups[g], targets[g]);
        if (!gpu) continue;
        // world B: SeekSystem::issuePaths2 (SeekSystem.cpp:12-29) against the components the slot owns
        auto& comps = seek->getComponents();
        for (int e : groups[g]) comps.at(seek->entity2compvec_ind_.at(e)).path_end = targets[g];
        B.pf->issuePaths2(comps, groups[g], seek->entity2compvec_ind_, targets[g]);
    }

    long mismatches = 0, compared = 0, uploaded = 0;
    int first_bad_frame = -1;
    int replans = 0;
    double t_cpu = 0, t_gpu = 0;
    // a later command (frame 25): half of the units that are still STANDING get a destination of their own. In world B this
    // is the only thing the blackboard shadow of the slot does not already know — exactly those entities are re-sent.
    std::vector<int> late;
    for (int i = 0; i < n_agents; ++i)
        if (A.shared->at(i).state == MoveState::STANDING && i % 2 == 0) late.push_back(i);
    const sf::Vector2f late_target = {1600.f, 900.f};   // (far from every unit: the arrival rule, an opt-in of the GPU path, must not fire)
    // Game::updateTriangulation in the middle of the run (frame 40): ONE more building, in the way of the units. World A gets it
    // through the reference's own code; world B's host-side CDT / Edges (the planner's) get it the same way, and the GPU
    // slot receives only the DIFFERENCE (rts_update_map_region): the triangle records that changed and the eight wall lines.
    long map_update_bytes = 0, map_full_bytes = 0;
    int map_changed_tris = 0, map_changed_lines = 0;
    for (int f = 0; f < frames; ++f) {
        if (f == 40) {
            int bx = -1, by = -1;
            for (int tries = 0; tries < 2000 && bx < 0; ++tries) {
                const int cx = 100 + rng() % 80, cy = 100 + rng() % 80;
                if (!blocked(cx - 7, cy - 7, cx + 7, cy + 7)) { bx = cx; by = cy; }
            }
            if (bx >= 0) {
                auto snapshot_lines = [&](RefWorld& w) {
                    std::vector<std::vector<Edgef>> lines;
                    for (int o = 0; o < 4; ++o)
                        for (auto& line : w.edges->edge_finders_[o]->edge_array2_) lines.emplace_back(line.begin(), line.end());
                    return lines;
                };
                const auto tris_before = B.cdt->triangles_;
                const auto lines_before = snapshot_lines(B);
                for (int k = 0; k < 2; ++k)
                    if (refh_add_building(W[k], bx, by, 8, 8, 1) != 0 || refh_finalize_map(W[k]) != 0) { printf("{\"error\": \"mid-run building\"}\n"); return 1; }
                if (gpu) {
                    auto& cdt = *B.cdt;
                    auto same_tri = [](const Triangle& a, const Triangle& b) {
                        for (int k = 0; k < 3; ++k)
                            if (a.verts[k].x != b.verts[k].x || a.verts[k].y != b.verts[k].y || a.neighbours[k] != b.neighbours[k] ||
                                a.is_constrained[k] != b.is_constrained[k]) return false;
                        return true;
                    };
                    std::vector<uint32_t> idx;
                    std::vector<RtsTriangle> rec;
                    for (size_t t = 0; t < cdt.triangles_.size(); ++t) {
                        if (t < tris_before.size() && same_tri(tris_before[t], cdt.triangles_[t])) continue;
                        RtsTriangle r{};
                        for (int k = 0; k < 3; ++k) {
                            r.vx[k] = cdt.triangles_[t].verts[k].x; r.vy[k] = cdt.triangles_[t].verts[k].y;
                            r.neighbours[k] = cdt.triangles_[t].neighbours[k]; r.is_constrained[k] = cdt.triangles_[t].is_constrained[k];
                        }
                        idx.push_back((uint32_t)t);
                        rec.push_back(r);
                    }
                    const auto lines_after = snapshot_lines(B);
                    std::vector<uint32_t> l_orient, l_index, l_count;
                    std::vector<RtsEdge> l_edges;
                    size_t flat = 0;
                    for (int o = 0; o < 4; ++o) {
                        const size_t n_lines = B.edges->edge_finders_[o]->edge_array2_.size();
                        for (size_t l = 0; l < n_lines; ++l, ++flat) {
                            const auto& a = lines_before[flat];
                            const auto& b = lines_after[flat];
                            bool same = a.size() == b.size();
                            for (size_t e = 0; same && e < a.size(); ++e)
                                same = a[e].from.x == b[e].from.x && a[e].from.y == b[e].from.y && a[e].t.x == b[e].t.x && a[e].t.y == b[e].t.y && a[e].l == b[e].l;
                            if (same) continue;
                            l_orient.push_back((uint32_t)o); l_index.push_back((uint32_t)l); l_count.push_back((uint32_t)b.size());
                            for (auto& s : b) l_edges.push_back({s.from.x, s.from.y, s.t.x, s.t.y, s.l});
                        }
                    }
                    if (rts_update_map_region(seek->h, (uint32_t)cdt.triangles_.size(), idx.data(), rec.data(), (uint32_t)idx.size(), l_orient.data(),
                                              l_index.data(), l_count.data(), l_edges.data(), (uint32_t)l_orient.size(), 0u) != RTS_OK) {
                        printf("{\"error\": \"rts_update_map_region: %s\"}\n", rts_last_error(seek->h));
                        return 2;
                    }
                    uint64_t last = 0, full = 0;
                    rts_map_upload_bytes(seek->h, &last, &full);
                    map_update_bytes = (long)last; map_full_bytes = (long)full;
                    map_changed_tris = (int)idx.size(); map_changed_lines = (int)l_orient.size();
                    // the patched cache against the table the reference's full updateCellGrid has just written (world B's host CDT)
                    std::vector<uint32_t> dev(cdt.cell2triangle_ind_.size());
                    if (rts_get_cell_grid(seek->h, dev.data()) != RTS_OK) { printf("{\"error\": \"rts_get_cell_grid\"}\n"); return 2; }
                    for (size_t c = 0; c < dev.size(); ++c) cell_grid_mismatches += dev[c] != (uint32_t)cdt.cell2triangle_ind_[c];
                }
            }
        }
        if (f == 25 && !late.empty()) {
            for (int e : late) { A.shared->at(e).state = MoveState::MOVING; B.shared->at(e).state = MoveState::MOVING; }
            A.ss->issuePaths2(late, late_target);
            if (gpu) {
                auto& comps = seek->getComponents();
                for (int e : late) comps.at(seek->entity2compvec_ind_.at(e)).path_end = late_target;
                B.pf->issuePaths2(comps, late, seek->entity2compvec_ind_, late_target);
            }
        }
        // ---------------- world A: ECSystem::update with the reference systems (as oracle/ref_harness.cpp)
        double t0 = omp_get_wtime();
        A.ps->updateSharedData(*A.shared, A.active);
        A.ss->updateSharedData(*A.shared, A.active);
        plan_pending(*A.pf, static_cast<ComponentArray<PathFinderComponent>&>(*A.ss->p_comps_).components_, A.ss->entity2compvec_ind_);
        A.ps->update();
        A.ss->update();
        A.ps->communicate(*A.shared);
        A.ss->communicate(*A.shared);
        A.ps->updateSharedData(*A.shared, A.active);
        A.ps->avoidWall();
        A.ps->communicateVelocities(*A.shared);
        std::vector<sf::Vector2f> velA(n_agents);
        for (auto ent : A.active) {
            auto& tc = A.shared->at(ent.ind).transform;
            velA[ent.ind] = tc.vel;
            tc.r += tc.vel * dt;
            tc.angle += tc.angle_vel;
            tc.vel *= 0.f;
        }
        t_cpu += omp_get_wtime() - t0;
        if (getenv("RTS_DEMO_TRACE")) fprintf(stderr, "frame %d: reference done\n", f);
        if (!gpu) continue;

        // ---------------- world B: the same loop over the GPU slots; ECS.cpp:92-95 removed
        t0 = omp_get_wtime();
        phys->updateSharedData(*B.shared, B.active);
        seek->updateSharedData(*B.shared, B.active);
        if (f > 0) uploaded += seek->uploaded_last_frame;   // (frame 0 sends the whole population once)
        {   // the planner part of SeekSystem::update (SeekSystem.cpp:65) stays on the host
            auto& comps = seek->getComponents();
            const size_t pending = B.pf->to_update_groups_.size();
            plan_pending(*B.pf, comps, seek->entity2compvec_ind_);
            if (pending) seek->markPathsDirty();
        }
        phys->update();
        seek->update();
        phys->communicate(*B.shared);
        seek->communicate(*B.shared);
        {   // SeekSystem.cpp:123-130: re-issue paths for the agents whose window was shifted
            auto& comps = seek->getComponents();
            for (int ci : seek->takeReplanRequests()) {
                std::vector<int> ent = {seek->compvec_ind2entity_ind_.at(ci)};
                B.pf->issuePaths2(comps, ent, seek->entity2compvec_ind_, comps.at(ci).path_end);
                ++replans;
            }
        }
        std::vector<sf::Vector2f> velB(n_agents);
        for (auto ent : B.active) {
            auto& tc = B.shared->at(ent.ind).transform;
            velB[ent.ind] = tc.vel;
            tc.r += tc.vel * dt;
            tc.angle += tc.angle_vel;
            tc.vel *= 0.f;
        }
        t_gpu += omp_get_wtime() - t0;

        // ---------------- compare the blackboards
        long bad = 0;
        for (int i = 0; i < n_agents; ++i) {
            const auto& a = A.shared->at(i);
            const auto& b = B.shared->at(i);
            const bool ok = same_bits(a.transform.r.x, b.transform.r.x) && same_bits(a.transform.r.y, b.transform.r.y) &&
                            same_bits(velA[i].x, velB[i].x) && same_bits(velA[i].y, velB[i].y) &&
                            same_bits(a.transform.angle, b.transform.angle) && same_bits(a.transform.angle_vel, b.transform.angle_vel) &&
                            a.state == b.state && same_bits(a.target.x, b.target.x) && same_bits(a.target.y, b.target.y);
            if (!ok && getenv("RTS_DEMO_TRACE") && mismatches + bad < 12)
                fprintf(stderr, "frame %d agent %d late %d: r %.9g %.9g | %.9g %.9g  vel %.9g %.9g | %.9g %.9g  angle %.9g | %.9g  angle_vel %.9g | %.9g  state %d | %d  target %.9g %.9g | %.9g %.9g\n",
                        f, i, (int)(std::find(late.begin(), late.end(), i) != late.end()), a.transform.r.x, a.transform.r.y, b.transform.r.x, b.transform.r.y,
                        velA[i].x, velA[i].y, velB[i].x, velB[i].y, a.transform.angle, b.transform.angle, a.transform.angle_vel, b.transform.angle_vel,
                        (int)a.state, (int)b.state, a.target.x, a.target.y, b.target.x, b.target.y);
            bad += !ok;
            ++compared;
        }
        if (bad && first_bad_frame < 0) first_bad_frame = f;
        mismatches += bad;
    }
    int moving = 0;
    for (int i = 0; i < n_agents; ++i) moving += A.shared->at(i).state == MoveState::MOVING;
    printf("{\"agents\": %d, \"buildings\": %d, \"frames\": %d, \"cell_grid_mismatches\": %ld, \"compared\": %ld, \"mismatches\": %ld, \"first_bad_frame\": %d, "
           "\"replans\": %d, \"moving_at_end\": %d, \"late_command\": %d, \"entities_resent_after_frame0\": %ld, \"map_update\": {\"changed_triangles\": %d, \"changed_wall_lines\": %d, \"bytes\": %ld, \"full_upload_bytes\": %ld}, "
           "\"cpu_ms_per_frame\": %.3f, \"gpu_adapter_ms_per_frame\": %.3f}\n",
           n_agents, n_buildings, frames, cell_grid_mismatches, compared, mismatches, first_bad_frame, replans, moving, (int)late.size(), uploaded,
           map_changed_tris, map_changed_lines, map_update_bytes, map_full_bytes,
           1e3 * t_cpu / frames, 1e3 * t_gpu / frames);
    return (mismatches || cell_grid_mismatches) ? 3 : 0;
}
