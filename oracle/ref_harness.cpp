// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Headless C-ABI harness around the *verbatim* RTSEngine translation units. It is
// compiled by oracle/Makefile together with the reference's own sources, read in place
// from /root/reference/src (nothing is copied into this repository), into
// oracle/_ref/libref_harness.so.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load it, and only as a checker
// or as the timed CPU baseline.
//
// What is restated here (because the owning TU needs OpenGL and cannot be compiled):
//   * the frame orchestration of ECSystem::update        (src/ECS.cpp:65-113)
//   * PhysicsSystem::communicate (private)                (src/Systems/PhysicsSystem.cpp:118-127)
//   * the UI default force multipliers                    (src/UI.cpp:124-131)
//   * UnitInitializer::initializeUnit component wiring    (src/UnitInitializer.cpp:19-80)
//   * the non-overlapping case of MapGrid::addBuildingEdgesToTriangulation
//     (src/MapGrid.cpp:493-541, 1095-1312) for octagonal buildings
// Everything else (neighbour search, forces, walls, seek, CDT, funnel planner) runs the
// reference's own compiled code.

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#include <algorithm>
#include <chrono>
#include <stdexcept>
#include <omp.h>

#define private public     // harness needs PhysicsSystem::communicate and PathFinder2::pathFromFunnel
#define protected public   // ... and System2::entity2compvec_ind_ (GCC class layout does not depend on access)
#include "Systems/PhysicsSystem.h"
#include "Systems/SeekSystem.h"
#include "PathFinding/PathFinder2.h"
#include "PathFinding/Triangulation.h"
#undef private
#undef protected
#include "Utils/Grid.h"
#include "Edges.h"

namespace {

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
    std::vector<sf::Vector2f> vel_prezero;   // blackboard velocity right before ECS.cpp:112 zeroes it
    int n_agents = 0;
    std::string last_error;
};

inline RefWorld* W(void* h) { return static_cast<RefWorld*>(h); }

// PathFinder2::updatePaths starts up to hardware_concurrency()-1 workers and indexes 12-entry per-thread arrays with
// the worker id (PathFinder2.h:161-162 N_MAX_THREADS, PathFinder2.cpp:90,132-137): on a host with more than 13 hardware
// threads the reference writes out of bounds. The harness therefore feeds the pending groups to the reference's own
// updatePaths at most 8 at a time (each group's path is independent of the others, so the result is the same).
void plan_pending(RefWorld* w) {
    auto& comps = static_cast<ComponentArray<PathFinderComponent>&>(*w->ss->p_comps_).components_;
    std::remove_reference_t<decltype(w->pf->to_update_groups_)> all;
    all.swap(w->pf->to_update_groups_);
    for (size_t i = 0; i < all.size(); i += 8) {
        w->pf->to_update_groups_.assign(all.begin() + i, all.begin() + std::min(i + 8, all.size()));
        w->pf->updatePaths(comps, w->ss->entity2compvec_ind_, 5000);
    }
}

}  // namespace

extern "C" {

// world with an n_cells_x × n_cells_y MapGrid of 5-unit cells and the app's triangle cache grid
// (cells/8 per side, 40 units; src/main.cpp:50).  NOTE: the physics/seek neighbour grids are
// hard-wired to Geometry::BOX = 2560² in the system constructors (PhysicsSystem.cpp:5-9).
void* refh_create(int n_cells_x, int n_cells_y) {
    auto* w = new RefWorld();
    w->n_cells = {n_cells_x, n_cells_y};
    w->tri_grid = std::make_unique<SearchGrid>(sf::Vector2i{n_cells_x / 8, n_cells_y / 8}, sf::Vector2f{40.f, 40.f});
    w->cdt = std::make_unique<Triangulation>(*w->tri_grid);
    w->cdt->createBoundaryAndSuperTriangle({n_cells_x * 5, n_cells_y * 5});
    w->edges = std::make_unique<Edges>(sf::Vector2i{n_cells_x, n_cells_y}, sf::Vector2f{5.f, 5.f});
    w->ps = std::make_unique<PhysicsSystem>(ComponentID::PHYSICS);
    w->ps->p_comps_ = std::make_shared<ComponentArray<PhysicsComponent>>();
    w->ps->p_map_ = w->edges.get();
    using M = PhysicsSystem::Multiplier;
    w->ps->force_multipliers[M::ALIGN] = 1.0f;     // src/UI.cpp:124-131
    w->ps->force_multipliers[M::SEEK] = 1.0f;
    w->ps->force_multipliers[M::SCATTER] = 0.1f;
    w->ps->force_multipliers[M::DECAY] = 1.0f;
    w->ps->force_multipliers[M::REPULSE] = 1.f;
    w->ps->force_multipliers[M::PUSH] = 0.1f;
    w->ps->force_multipliers[M::VELOCITY] = 2.0f;
    w->ss = std::make_unique<SeekSystem>(ComponentID::PATHFINDING);
    w->ss->p_comps_ = std::make_shared<ComponentArray<PathFinderComponent>>();
    w->pf = std::make_unique<PathFinder2>(w->cdt.get());
    w->ss->p_pathfinder_ = w->pf.get();
    w->ss->p_cdt_ = w->cdt.get();
    w->shared = std::make_unique<std::array<SharedData, N_MAX_ENTITIES>>();
    return w;
}

void refh_destroy(void* h) { delete W(h); }

const char* refh_last_error(void* h) { return W(h)->last_error.c_str(); }

void refh_set_multiplier(void* h, int which, float value) {
    W(h)->ps->force_multipliers[static_cast<PhysicsSystem::Multiplier>(which)] = value;
}

// Octagonal building, n×m cells, corner size c, centre cell (cx,cy).  Restates the non-overlapping
// branch of MapGrid::addBuildingEdgesToTriangulation (MapGrid.cpp:1095-1312) on top of the
// reference's own Triangulation / EdgesFinder1D.  Returns 0, or -1 when the CDT rejected it.
int refh_add_building(void* h, int cx, int cy, int n, int m, int c) {
    auto* w = W(h);
    try {
        auto& cdt = *w->cdt;
        auto& E = *w->edges;
        const int ulx = cx - n / 2, uly = cy - m / 2;
        const int hl = n - 2 * c, vl = m - 2 * c;
        // Building::intitializeEdges (MapGrid.cpp:493-541): eight outline vertices, clockwise (y down)
        Vertex v[8] = {{ulx + c, uly},          {ulx + n - c, uly},      {ulx + n, uly + c},      {ulx + n, uly + c + vl},
                       {ulx + c + hl, uly + m}, {ulx + c, uly + m},      {ulx, uly + c + vl},     {ulx, uly + c}};
        for (auto& p : v) { p.x *= 5; p.y *= 5; }
        const int ny = w->n_cells.y;
        auto cellOf = [&](Vertex p) { return sf::Vector2i{static_cast<int>((p.x + 2) / 5.f), static_cast<int>((p.y + 2) / 5.f)}; };
        // line indices (MapGrid.cpp:1107-1134)
        const int line_axis[4] = {uly, ulx + n, uly + m, ulx};
        const auto c1 = cellOf(v[1]), c3 = cellOf(v[3]), c5 = cellOf(v[5]), c7 = cellOf(v[7]);
        const int line_diag[4] = {c1.x - c1.y + ny - 1, c3.x + c3.y, c5.x - c5.y + ny - 1, c7.x + c7.y - 1};

        VertInd vi[8];
        // insertion order of the axis-aligned pairs: (0,1) (2,3) (5,4) (7,6)   (MapGrid.cpp:1183-1190 with `inv`)
        const int order[8] = {0, 1, 2, 3, 5, 4, 7, 6};
        for (int k = 0; k < 8; ++k) {
            const int id = order[k];
            auto data = cdt.insertVertexAndGetData(v[id], true);
            if (data.overlapping_vertex != (VertInd)-1 || data.overlapping_edge.from != (VertInd)-1) {
                w->last_error = "building touches an existing vertex/edge (harness supports separated buildings only)";
                return -1;
            }
            vi[id] = cdt.vertices_.size() - 1;
        }
        auto pushEdge = [&](int orient, int line, const Edgef& e) {
            E.edges2_.push_back({0u, 0u});   // insert3 swaps indices in edges2_[idx]; keep the index valid
            E.edges_.push_back(e);
            E.edge_finders_.at(orient)->insert3(e, E.edges2_.size() - 1, line);
        };
        // axis-aligned walls: i=0,1 run v[2i]→v[2i+1]; i=2,3 run the other way (MapGrid.cpp:1258-1268),
        // so that (t.y,-t.x) is always the outward normal.
        pushEdge(0, line_axis[0], Edgef(v[0], v[1]));
        pushEdge(2, line_axis[1], Edgef(v[2], v[3]));
        pushEdge(0, line_axis[2], Edgef(v[4], v[5]));
        pushEdge(2, line_axis[3], Edgef(v[6], v[7]));
        for (int i = 0; i < 4; ++i) {
            auto& line = E.edge_finders_.at(2 * (i % 2))->edge_array2_.at(line_axis[i]);
            if (i % 2 == 0)
                std::sort(line.begin(), line.end(), [](const Edgef& a, const Edgef& b) {
                    return std::min({a.from.x, a.to().x}) < std::min({b.from.x, b.to().x});
                });
            else
                std::sort(line.begin(), line.end(), [](const Edgef& a, const Edgef& b) { return a.from.y < b.from.y; });
        }
        // diagonal walls v[2i+1]→v[2i+2], finder 1 for i even, 3 for i odd (MapGrid.cpp:1291-1303)
        for (int i = 0; i < 4; ++i) {
            const int orient = 2 * (i % 2) + 1;
            pushEdge(orient, line_diag[i], Edgef(v[2 * i + 1], v[(2 * i + 2) % 8]));
            auto& line = E.edge_finders_.at(orient)->edge_array2_.at(line_diag[i]);
            std::sort(line.begin(), line.end(), [](const Edgef& a, const Edgef& b) { return a.from.x < b.from.x; });
        }
        // constraints, same order as new_constraints (MapGrid.cpp:1252-1310)
        cdt.insertConstraint({vi[0], vi[1]});
        cdt.insertConstraint({vi[2], vi[3]});
        cdt.insertConstraint({vi[5], vi[4]});
        cdt.insertConstraint({vi[7], vi[6]});
        for (int i = 0; i < 4; ++i) cdt.insertConstraint({vi[2 * i + 1], vi[(2 * i + 2) % 8]});
    } catch (std::exception& e) {
        w->last_error = e.what();
        return -1;
    }
    return 0;
}

// PathFinder2::update(): triangle widths, reduced graph, Triangulation::updateCellGrid  (PathFinder2.cpp:63-83)
int refh_finalize_map(void* h) {
    try {
        W(h)->pf->update();
    } catch (std::exception& e) {
        W(h)->last_error = e.what();
        return -1;
    }
    return 0;
}
// only the cell→triangle cache (no reduced graph): enough for findTriangle
void refh_update_cell_grid(void* h) { W(h)->cdt->updateCellGrid(); }
// the same with Triangulation::last_found_tri_ind_ (where the first cell's walk starts) set by the caller
void refh_rebuild_cell_grid(void* h, int last_found) {
    W(h)->cdt->last_found_tri_ind_ = last_found;
    W(h)->cdt->updateCellGrid();
}
int refh_last_found(void* h) { return (int)W(h)->cdt->last_found_tri_ind_; }
int refh_consistent(void* h) { return W(h)->cdt->triangulationIsConsistent() ? 1 : 0; }

// UnitInitializer::initializeUnit (UnitInitializer.cpp:19-80): pc.mass/radius from the unit type, pfc.radius = pc.radius.
int refh_add_agent(void* h, float x, float y, float angle, float radius, float mass, int player, int state) {
    auto* w = W(h);
    const int e = w->n_agents;
    if (e >= N_MAX_NAVIGABLE_BOIDS) { w->last_error = "N_MAX_NAVIGABLE_BOIDS"; return -1; }
    TransformComponent tc;
    tc.r = {x, y};
    tc.vel = {0, 0};
    tc.angle = angle;
    tc.angle_vel = 0;
    PhysicsComponent pc;
    pc.transform = tc;
    pc.mass = mass;
    pc.radius = radius;
    pc.player_ind = player;
    PathFinderComponent pfc;
    pfc.transform = tc;
    pfc.radius = pc.radius;
    Entity ent(0);
    ent.ind = e;
    SharedData sd;
    sd.transform = tc;
    sd.state = static_cast<MoveState>(state);
    w->shared->at(e) = sd;
    w->ps->addComponent(ent, pc);
    w->ss->addComponent(ent, pfc);
    w->active.push_back(ent);
    w->n_agents++;
    return e;
}

int refh_n_agents(void* h) { return W(h)->n_agents; }

// right-click move command: ECSystem::setMoveState(MOVING) + SeekSystem::issuePaths2   (Game.cpp:211-231)
int refh_issue_paths(void* h, const int* ids, int n, float tx, float ty) {
    auto* w = W(h);
    try {
        std::vector<int> sel(ids, ids + n);
        for (int e : sel) w->shared->at(e).state = MoveState::MOVING;
        w->ss->issuePaths2(sel, {tx, ty});
    } catch (std::exception& e) {
        w->last_error = e.what();
        return -1;
    }
    return 0;
}

// run the planner now (outside a step) so the windows can be read before stepping
int refh_plan_now(void* h) {
    auto* w = W(h);
    try {
        plan_pending(w);
    } catch (std::exception& e) {
        w->last_error = e.what();
        return -1;
    }
    return 0;
}

// One frame of ECSystem::update (ECS.cpp:65-113) restricted to the Physics and Seek systems,
// in ComponentID order PHYSICS(4) then PATHFINDING(5).
int refh_step(void* h, int n_steps) {
    auto* w = W(h);
    auto& shared = *w->shared;
    try {
        for (int s = 0; s < n_steps; ++s) {
            w->ps->updateSharedData(shared, w->active);
            w->ss->updateSharedData(shared, w->active);
            plan_pending(w);   // what SeekSystem::update does first (SeekSystem.cpp:65), in safe portions
            w->ps->update();
            w->ss->update();
            w->ps->communicate(shared);
            w->ss->communicate(shared);
            w->ps->updateSharedData(shared, w->active);
            w->ps->avoidWall();
            w->ps->communicateVelocities(shared);
            w->vel_prezero.resize(w->n_agents);
            for (auto ent : w->active) {
                auto& tc = shared.at(ent.ind).transform;
                w->vel_prezero[ent.ind] = tc.vel;
                tc.r += tc.vel * dt;
                tc.angle += tc.angle_vel;
                tc.vel *= 0.f;
            }
        }
    } catch (std::exception& e) {
        w->last_error = e.what();
        return -1;
    }
    return 0;
}

// blackboard + per-system state after the last frame. Any pointer may be null.
void refh_get_state(void* h, float* r, float* vel, float* inertia, float* angle, float* angle_vel, int* state) {
    auto* w = W(h);
    auto& pcs = w->ps->getComponents();
    for (int i = 0; i < w->n_agents; ++i) {
        const auto& sd = w->shared->at(i);
        if (r) { r[2 * i] = sd.transform.r.x; r[2 * i + 1] = sd.transform.r.y; }
        if (vel && i < (int)w->vel_prezero.size()) { vel[2 * i] = w->vel_prezero[i].x; vel[2 * i + 1] = w->vel_prezero[i].y; }
        if (inertia) { inertia[2 * i] = pcs[i].inertia_vel.x; inertia[2 * i + 1] = pcs[i].inertia_vel.y; }
        if (angle) angle[i] = sd.transform.angle;
        if (angle_vel) angle_vel[i] = sd.transform.angle_vel;
        if (state) state[i] = static_cast<int>(sd.state);
    }
}

void refh_set_state(void* h, const float* r, const float* inertia, const float* angle, const float* angle_vel, const int* state) {
    auto* w = W(h);
    auto& pcs = w->ps->getComponents();
    for (int i = 0; i < w->n_agents; ++i) {
        auto& sd = w->shared->at(i);
        if (r) sd.transform.r = {r[2 * i], r[2 * i + 1]};
        if (inertia) pcs[i].inertia_vel = {inertia[2 * i], inertia[2 * i + 1]};
        if (angle) sd.transform.angle = angle[i];
        if (angle_vel) sd.transform.angle_vel = angle_vel[i];
        if (state) sd.state = static_cast<MoveState>(state[i]);
    }
}

void refh_set_phys_params(void* h, int i, float lambda, float max_speed) {
    auto& pc = W(h)->ps->getComponents().at(i);
    pc.lambda = lambda;
    pc.max_speed = max_speed;
}
void refh_set_seek_params(void* h, int i, float max_speed, float turn_rate, float desired_angle) {
    auto& c = static_cast<ComponentArray<PathFinderComponent>&>(*W(h)->ss->p_comps_).components_.at(i);
    c.max_speed = max_speed;
    c.turn_rate = turn_rate;
    c.desired_angle = desired_angle;
}

// seek window, 16 floats per agent:
// r_next(2) r_next_next(2) | portal_next.from(2) .t(2) | portal_nn.from(2) .t(2) | portal_next.l portal_nn.l path_end(2)
void refh_get_seek_window(void* h, float* win) {
    auto* w = W(h);
    auto& comps = static_cast<ComponentArray<PathFinderComponent>&>(*w->ss->p_comps_).components_;
    for (int i = 0; i < w->n_agents; ++i) {
        const auto& c = comps[i];
        float* o = win + 16 * i;
        o[0] = c.r_next.x; o[1] = c.r_next.y; o[2] = c.r_next_next.x; o[3] = c.r_next_next.y;
        o[4] = c.portal_next.from.x; o[5] = c.portal_next.from.y; o[6] = c.portal_next.t.x; o[7] = c.portal_next.t.y;
        o[8] = c.portal_next_next.from.x; o[9] = c.portal_next_next.from.y; o[10] = c.portal_next_next.t.x; o[11] = c.portal_next_next.t.y;
        o[12] = c.portal_next.l; o[13] = c.portal_next_next.l; o[14] = c.path_end.x; o[15] = c.path_end.y;
    }
}
void refh_set_seek_window(void* h, const float* win) {
    auto* w = W(h);
    auto& comps = static_cast<ComponentArray<PathFinderComponent>&>(*w->ss->p_comps_).components_;
    for (int i = 0; i < w->n_agents; ++i) {
        auto& c = comps[i];
        const float* o = win + 16 * i;
        c.r_next = {o[0], o[1]}; c.r_next_next = {o[2], o[3]};
        c.portal_next.from = {o[4], o[5]}; c.portal_next.t = {o[6], o[7]};
        c.portal_next_next.from = {o[8], o[9]}; c.portal_next_next.t = {o[10], o[11]};
        c.portal_next.l = o[12]; c.portal_next_next.l = o[13]; c.path_end = {o[14], o[15]};
    }
}
// number of re-plan requests the last frame queued (SeekSystem.cpp:294-303); they are consumed inside update()
int refh_group_of(void* h, int i) { return W(h)->ss->commander_groups_.entity2command_group.at(i); }

// ---- tables ---------------------------------------------------------------------------------
int refh_n_triangles(void* h) { return (int)W(h)->cdt->triangles_.size(); }
// verts: 6 int per triangle; nbrs: 3 u32; constrained: 3 u8
void refh_get_triangles(void* h, int32_t* verts, uint32_t* nbrs, uint8_t* constrained) {
    const auto& tris = W(h)->cdt->triangles_;
    for (size_t t = 0; t < tris.size(); ++t)
        for (int k = 0; k < 3; ++k) {
            verts[6 * t + 2 * k] = tris[t].verts[k].x;
            verts[6 * t + 2 * k + 1] = tris[t].verts[k].y;
            nbrs[3 * t + k] = tris[t].neighbours[k];
            if (constrained) constrained[3 * t + k] = tris[t].is_constrained[k];
        }
}
int refh_tri_grid_cells(void* h) { return (int)W(h)->cdt->cell2triangle_ind_.size(); }
void refh_get_cell2tri(void* h, uint32_t* out) {
    const auto& c = W(h)->cdt->cell2triangle_ind_;
    std::memcpy(out, c.data(), c.size() * sizeof(uint32_t));
}
int refh_n_lines(void* h, int orient) { return (int)W(h)->edges->edge_finders_.at(orient)->edge_array2_.size(); }
int refh_n_edges(void* h, int orient) {
    int n = 0;
    for (auto& l : W(h)->edges->edge_finders_.at(orient)->edge_array2_) n += (int)l.size();
    return n;
}
// CSR of one orientation: offsets[n_lines+1], edges 5 floats each {from.x from.y t.x t.y l}
void refh_get_edges(void* h, int orient, uint32_t* offsets, float* edges) {
    const auto& lines = W(h)->edges->edge_finders_.at(orient)->edge_array2_;
    uint32_t k = 0;
    for (size_t l = 0; l < lines.size(); ++l) {
        offsets[l] = k;
        for (const auto& e : lines[l]) {
            float* o = edges + 5 * k++;
            o[0] = e.from.x; o[1] = e.from.y; o[2] = e.t.x; o[3] = e.t.y; o[4] = e.l;
        }
    }
    offsets[lines.size()] = k;
}

// ---- point queries (known-answer / parity helpers) ------------------------------------------
// Triangulation::findTriangle(Vector2f,false)   (Triangulation.cpp:1588-1590 → 1497-1567)
void refh_find_triangles(void* h, const float* xy, int n, uint32_t* out) {
    auto& cdt = *W(h)->cdt;
    for (int i = 0; i < n; ++i) out[i] = cdt.findTriangle(sf::Vector2f{xy[2 * i], xy[2 * i + 1]}, false);
}
uint32_t refh_find_triangle_int(void* h, int x, int y) { return W(h)->cdt->findTriangle(Vertex{x, y}, false); }
void refh_insert_vertex(void* h, int x, int y, int from_last) { W(h)->cdt->insertVertex(Vertex{x, y}, from_last != 0); }

// Grid::coordToCell   (Grid.cpp:18-24)
void refh_coord_to_cell(int nx, int ny, float cs, const float* xy, int n, uint64_t* out) {
    Grid g({nx, ny}, {cs, cs});
    for (int i = 0; i < n; ++i) out[i] = g.coordToCell(sf::Vector2f{xy[2 * i], xy[2 * i + 1]});
}
int refh_physics_grid_nx(void* h) { return W(h)->ps->p_ns_->s_grid_->n_cells_.x; }

// Edges::calcContactEdgesO   (Edges.cpp:245-296): writes up to `cap` edges (5 floats each) in visit order, returns count
int refh_contact_edges(void* h, float x, float y, float radius, float* out, int cap) {
    auto walls = W(h)->edges->calcContactEdgesO({x, y}, radius);
    int k = 0;
    for (int o = 0; o < 4; ++o)
        for (const auto& e : walls[o]) {
            if (k < cap) {
                float* p = out + 5 * k;
                p[0] = e.from.x; p[1] = e.from.y; p[2] = e.t.x; p[3] = e.t.y; p[4] = e.l;
            }
            ++k;
        }
    return k;
}

// physics neighbour lists of the last p_ns_->update: neighbour component indices in visit order
int refh_physics_neighbours(void* h, int i, int* out, int cap) {
    auto& ns = *W(h)->ps->p_ns_;
    const int n = ns.last_i.at(i);
    const auto& d = ns.getInteractionData(i);
    for (int k = 0; k < n && k < cap; ++k) out[k] = d[k].second;
    return n;
}
// run only the physics neighbour search on the current blackboard positions
void refh_physics_ns_update(void* h) {
    auto* w = W(h);
    w->ps->updateSharedData(*w->shared, w->active);
    w->ps->p_ns_->update(w->ps->getComponents());
}

// Stand-alone seek neighbour searcher for the reference's known-answer test
// (src/Tests/test_neighbour_search.cc:8-66): returns neighbour count of agent `query`, first neighbour in *first.
int refh_seek_ns_test(float box, float cell, const float* xy, const float* radius, int n, int query, int* first) {
    NeighbourSearcherContext<PathFinderComponent, int> ns({box, box}, cell);
    ns.setStrategy(NS_STRATEGY::BASIC);
    std::vector<PathFinderComponent> comps(n);
    for (int i = 0; i < n; ++i) {
        comps[i].radius = radius[i];
        comps[i].transform.r = {xy[2 * i], xy[2 * i + 1]};
    }
    ns.update(comps);
    const int cnt = ns.last_i.at(query);
    if (first) *first = ns.getInteractionData(query).at(0);
    return cnt;
}
int refh_seek_ns_full(float box, float cell, const float* xy, const float* radius, int n, float qx, float qy, float rmax) {
    NeighbourSearcherContext<PathFinderComponent, int> ns({box, box}, cell);
    ns.setStrategy(NS_STRATEGY::BASIC);
    std::vector<PathFinderComponent> comps(n);
    for (int i = 0; i < n; ++i) {
        comps[i].radius = radius[i];
        comps[i].transform.r = {xy[2 * i], xy[2 * i + 1]};
    }
    ns.update(comps);
    return (int)ns.getNeighboursIndsFull({qx, qy}, rmax).size();
}

// ... and the index list itself, in the reference's visit order (pins the oracle's restatement of getNeighboursIndsFull)
int refh_ns_full_list(float box, float cell, const float* xy, int n, float qx, float qy, float rmax, int* out, int cap) {
    NeighbourSearcherContext<PathFinderComponent, int> ns({box, box}, cell);
    ns.setStrategy(NS_STRATEGY::BASIC);
    std::vector<PathFinderComponent> comps(n);
    for (int i = 0; i < n; ++i) comps[i].transform.r = {xy[2 * i], xy[2 * i + 1]};
    ns.update(comps);
    const auto inds = ns.getNeighboursIndsFull({qx, qy}, rmax);
    for (int k = 0; k < (int)inds.size() && k < cap; ++k) out[k] = inds[k];
    return (int)inds.size();
}

// The Attack strategy's nearest-enemy search as the reference runs it (AttackSystem.cpp:24: a context of 20*RHARD = 60-unit cells,
// NeighbourSearcherStrategy.cpp:125-177): out[i] = the enemy index the reference reports for agent i, or -1 when it reports none.
int refh_attack_ns(float box, float cell, const float* xy, const unsigned char* player, int n, int* out) {
    NeighbourSearcherContext<AttackComponent, int> ns({box, box}, cell);
    ns.setStrategy(NS_STRATEGY::ATTACK);
    std::vector<AttackComponent> comps(n);
    for (int i = 0; i < n; ++i) {
        comps[i].transform.r = {xy[2 * i], xy[2 * i + 1]};
        comps[i].player_ind = player[i];
    }
    ns.update(comps);
    for (int i = 0; i < n; ++i) out[i] = ns.last_i.at(i) > 0 ? ns.getInteractionData(i).at(0) : -1;
    return 0;
}

// full funnel path of one planner query (PathFinder2.cpp:388-438): waypoints (2 floats) and portals (5 floats)
int refh_plan_path(void* h, float sx, float sy, float ex, float ey, float radius, float* path, float* portals, int cap) {
    auto* w = W(h);
    try {
        Funnel funnel;
        w->pf->findSubOptimalPathCenters({sx, sy}, {ex, ey}, radius, funnel, 0);
        funnel.push_back({{sx, sy}, {sx, sy}});
        std::reverse(funnel.begin(), funnel.end());
        funnel.push_back({{ex, ey}, {ex, ey}});
        auto pap = w->pf->pathFromFunnel({sx, sy}, {ex, ey}, radius, funnel);
        const int n = (int)pap.path.size();
        for (int k = 0; k < n && k < cap; ++k) {
            path[2 * k] = pap.path[k].x; path[2 * k + 1] = pap.path[k].y;
            const auto& e = pap.portals.at(k);
            float* p = portals + 5 * k;
            p[0] = e.from.x; p[1] = e.from.y; p[2] = e.t.x; p[3] = e.t.y; p[4] = e.l;
        }
        return n;
    } catch (std::exception& e) {
        w->last_error = e.what();
        return -1;
    }
}

int refh_omp_threads() { return omp_get_max_threads(); }
void refh_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }

}  // extern "C"
