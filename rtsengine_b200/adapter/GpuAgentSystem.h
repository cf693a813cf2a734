// Drop-in adapter: the fused B200 agent update behind the reference's own ECS plugin interface.
//
// Compile this header INSIDE the reference tree (it includes the reference's ECS.h / Components.h) and link
// librts_b200.so.  It provides two `System2` subclasses (reference: src/ECS.h:147-234) that take the slots of
// PhysicsSystem and SeekSystem in ECSystem::systems2_ (src/ECS.cpp:16-17):
//
//   GpuPhysicsSlot  : System2   id = PHYSICS      holds the PhysicsComponents; update() is empty
//   GpuSeekSlot     : System2   id = PATHFINDING  holds the PathFinderComponents (so PathFinder2 and
//                                                  SeekSystem::issuePaths2-style commands keep working) and runs
//                                                  the whole fused frame on the GPU in update()
//
// Frame protocol (ECSystem::update, src/ECS.cpp:65-113):
//   updateSharedData : the entities whose blackboard entry differs from what the device already holds (a command changed
//                      the state, the game moved a unit) -> rts_update_agents_indexed                     (H2D, few records)
//   update           : windows the planner rewrote -> rts_upload_paths ; rts_step(1)
//   communicate      : rts_download -> SharedData.transform.vel = merged, wall-projected velocity of the frame,
//                      .angle_vel, .state, .target (21 B per agent + 8 B of re-plan flags / waypoint index).  The loop's own
//                      integration (ECS.cpp:98-113) then computes r += vel*dt, which is bit-identical to the position the
//                      GPU integrated — so the device state IS the blackboard and nothing has to travel back up. The slot
//                      keeps a shadow of what the blackboard will hold after that integration; updateSharedData compares.
// The second wall pass that ECSystem::update runs on the PhysicsSystem (ECS.cpp:92-95) is already part of the
// fused frame, so those three lines are removed when the adapter is installed.
//
// Limits: N <= N_MAX_ENTITIES (the reference's own std::array signature).  Errors: a non-zero status of the C ABI
// becomes std::runtime_error, the reference's own error convention (NeighbourSearcherContext.cpp:46).
#pragma once

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "ECS.h"              // reference: System2, SharedData, Entity, ComponentArray
#include "rts_gpu.h"

struct GpuPhysicsSlot : System2 {
    explicit GpuPhysicsSlot(ComponentID id) : System2(id) {}
    void update() override {}
    void updateSharedData(const std::array<SharedData, N_MAX_ENTITIES>&, const std::vector<Entity>&) override {}
    void communicate(std::array<SharedData, N_MAX_ENTITIES>&) const override {}
    std::vector<PhysicsComponent>& getComponents() { return static_cast<ComponentArray<PhysicsComponent>&>(*p_comps_).components_; }
    int entityOf(int comp_ind) const { return compvec_ind2entity_ind_.at(comp_ind); }
};

struct GpuSeekSlot : System2 {
    RtsHandle h = nullptr;
    GpuPhysicsSlot* physics = nullptr;   // same component order as this slot (both filled by ECSystem::initializeEntity)
    bool population_dirty = true;        // set after addComponent / remove
    bool paths_dirty = true;             // set after the planner rewrote windows (PathFinder2::setPathOfAgent)
    // command group of each component (CommandGroupManager2::entity2command_group, CommandGroupManager.h:130); the host
    // copies it here after addGroup / removeFromGroup. Empty: every agent is its own group (no arrival cascade).
    std::vector<int32_t> command_group;
    bool mirror_inertia = true;          // also copy PhysicsComponent::inertia_vel back every frame (8 B per agent; nothing on the GPU path reads it)
    uint32_t uploaded_last_frame = 0;    // entities the last updateSharedData had to send (diagnostic)

    GpuSeekSlot(ComponentID id, const RtsParams& params, int device = 0) : System2(id) {
        check(rts_create(&params, device, &h), "rts_create");
    }
    // System2 has no virtual destructor (ECS.h:147): the owner must destroy the slot through its concrete type
    ~GpuSeekSlot() { if (h) rts_destroy(h); }

    std::vector<PathFinderComponent>& getComponents() { return static_cast<ComponentArray<PathFinderComponent>&>(*p_comps_).components_; }

    // Edges / Triangulation flattened by the caller after Game::updateTriangulation (Game.cpp:38-48)
    void uploadMap(const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri, const RtsWallTable& walls) {
        check(rts_upload_map(h, tris, n_tris, cell2tri, &walls), "rts_upload_map");
    }

    void updateSharedData(const std::array<SharedData, N_MAX_ENTITIES>& data, const std::vector<Entity>& active) override {
        auto& seek = getComponents();
        const uint32_t n = (uint32_t)seek.size();
        const bool population_was_dirty = population_dirty;
        const bool full = population_dirty || shadow_r_.size() != 2 * (size_t)n;
        if (population_dirty) uploadPopulation(data);
        if (full) {
            shadow_r_.assign(2 * (size_t)n, 0.f); shadow_angle_.assign(n, 0.f); shadow_angle_vel_.assign(n, 0.f); shadow_state_.assign(n, 0);
        }
        dirty_.clear(); d_r_.clear(); d_angle_.clear(); d_angle_vel_.clear(); d_state_.clear();
        for (const auto ent : active) {
            const int i = entity2compvec_ind_.at(ent.ind);
            if (i < 0) continue;
            const auto& sd = data.at(ent.ind);
            const uint8_t st = (uint8_t)sd.state;
            // bitwise comparison with the shadow (a NaN position never equals itself: it is simply re-sent)
            const bool same = !full && sd.transform.r.x == shadow_r_[2 * i] && sd.transform.r.y == shadow_r_[2 * i + 1] &&
                              sd.transform.angle == shadow_angle_[i] && sd.transform.angle_vel == shadow_angle_vel_[i] && st == shadow_state_[i];
            if (!same) {
                if (!population_was_dirty) {   // (a full upload has just sent everything)
                    dirty_.push_back((uint32_t)i);
                    d_r_.push_back(sd.transform.r.x); d_r_.push_back(sd.transform.r.y);
                    d_angle_.push_back(sd.transform.angle); d_angle_vel_.push_back(sd.transform.angle_vel); d_state_.push_back(st);
                }
                shadow_r_[2 * i] = sd.transform.r.x; shadow_r_[2 * i + 1] = sd.transform.r.y;
                shadow_angle_[i] = sd.transform.angle; shadow_angle_vel_[i] = sd.transform.angle_vel; shadow_state_[i] = st;
            }
            seek[i].transform = sd.transform;   // the host-side planner reads positions from the components
            seek[i].state = sd.state;
        }
        uploaded_last_frame = (uint32_t)dirty_.size();
        if (!dirty_.empty()) {
            RtsAgentView v{};
            v.r = d_r_.data(); v.angle = d_angle_.data(); v.angle_vel = d_angle_vel_.data(); v.state = d_state_.data();
            check(rts_update_agents_indexed(h, dirty_.data(), (uint32_t)dirty_.size(), &v), "rts_update_agents_indexed");
            check(rts_sync(h), "rts_sync");   // (the staging copies read the vectors above)
        }
    }

    void update() override {
        if (paths_dirty) uploadWindows();
        check(rts_step(h, 1), "rts_step");
    }

    void communicate(std::array<SharedData, N_MAX_ENTITIES>& data) const override {
        auto* self = const_cast<GpuSeekSlot*>(this);
        auto& seek = self->getComponents();
        const uint32_t n = (uint32_t)seek.size();
        self->vel_.resize(2 * n); self->r_next_.resize(2 * n); self->flags_.resize(n); self->angle_vel_.resize(n); self->state_.resize(n);
        self->waypoint_.resize(n);
        RtsAgentView v{};
        v.vel = self->vel_.data(); v.angle_vel = self->angle_vel_.data(); v.state = self->state_.data();
        v.r_next = self->r_next_.data(); v.flags = self->flags_.data();
        v.waypoint = self->waypoint_.data();
        if (physics && mirror_inertia) { self->inertia_.resize(2 * n); v.inertia = self->inertia_.data(); }
        self->check(rts_download(h, &v), "rts_download");
        const float dt_frame = dt;   // ECS.h:18, the step the loop integrates with
        for (uint32_t i = 0; i < n; ++i) {
            auto& sd = data.at(compvec_ind2entity_ind_.at(i));
            sd.target = {r_next_[2 * i], r_next_[2 * i + 1]};                 // SeekSystem.cpp:330
            sd.state = static_cast<MoveState>(state_[i]);                     // :331
            sd.transform.vel = {vel_[2 * i], vel_[2 * i + 1]};                // PhysicsSystem.cpp:118-127 + SeekSystem.cpp:337
            sd.transform.angle_vel = angle_vel_[i];                           // :339
            // what the blackboard will hold once the loop has integrated (ECS.cpp:110-111) = what the device holds now
            if (self->shadow_r_.size() == 2 * (size_t)n) {
                sf::Vector2f r = sd.transform.r;
                r += sd.transform.vel * dt_frame;
                float angle = sd.transform.angle;
                angle += sd.transform.angle_vel;
                self->shadow_r_[2 * i] = r.x; self->shadow_r_[2 * i + 1] = r.y;
                self->shadow_angle_[i] = angle; self->shadow_angle_vel_[i] = angle_vel_[i]; self->shadow_state_[i] = state_[i];
            }
            auto& c = seek[i];
            c.r_next = sd.target;
            if (waypoint_[i] > 0) { c.portal_next = c.portal_next_next; }     // updateTarget shifted the window (:290-291)
            if (flags_[i] & RTS_FLAG_REPLAN) { c.needs_update = true; self->replan_.push_back((int)i); }   // :294-303
            if (physics && mirror_inertia) physics->getComponents()[i].inertia_vel = {inertia_[2 * i], inertia_[2 * i + 1]};
        }
    }

    // component indices whose window was shifted and that want a new plan (SeekSystem::comp_inds_to_update_)
    std::vector<int> takeReplanRequests() { std::vector<int> out; out.swap(replan_); return out; }

    void markPopulationDirty() { population_dirty = true; paths_dirty = true; }
    void markPathsDirty() { paths_dirty = true; }

private:
    std::vector<float> angle_vel_, vel_, r_next_, inertia_;
    std::vector<uint8_t> state_;
    // shadow of the blackboard as the device holds it, and the per-frame dirty list
    std::vector<float> shadow_r_, shadow_angle_, shadow_angle_vel_, d_r_, d_angle_, d_angle_vel_;
    std::vector<uint8_t> shadow_state_, d_state_;
    std::vector<uint32_t> dirty_;
    std::vector<uint32_t> flags_;
    std::vector<int32_t> waypoint_;
    std::vector<int> replan_;

    void check(int32_t rc, const char* what) const {
        if (rc != RTS_OK) throw std::runtime_error(std::string(what) + ": " + rts_last_error(h));
    }

    // System2::addComponent / remove changed the dense arrays: re-send everything (rare)
    void uploadPopulation(const std::array<SharedData, N_MAX_ENTITIES>& data) {
        auto& seek = getComponents();
        const uint32_t n = (uint32_t)seek.size();
        std::vector<float> r(2 * n), inertia(2 * n), angle(n), angle_vel(n), radius(n), mass(n);
        std::vector<uint8_t> state(n), player(n);
        for (uint32_t i = 0; i < n; ++i) {
            const auto& sd = data.at(compvec_ind2entity_ind_.at(i));
            r[2 * i] = sd.transform.r.x; r[2 * i + 1] = sd.transform.r.y;
            angle[i] = sd.transform.angle; angle_vel[i] = sd.transform.angle_vel;
            state[i] = (uint8_t)sd.state;
            radius[i] = seek[i].radius; mass[i] = 1.f;
            if (physics) {
                const auto& pc = physics->getComponents().at(i);
                radius[i] = pc.radius; mass[i] = pc.mass; player[i] = (uint8_t)pc.player_ind;
                inertia[2 * i] = pc.inertia_vel.x; inertia[2 * i + 1] = pc.inertia_vel.y;
            }
        }
        RtsAgentView v{};
        v.r = r.data(); v.inertia = inertia.data(); v.angle = angle.data(); v.angle_vel = angle_vel.data();
        v.radius = radius.data(); v.mass = mass.data(); v.state = state.data(); v.player = player.data();
        check(rts_set_agents(h, n, &v), "rts_set_agents");
        population_dirty = false;
        paths_dirty = true;
    }

    // one 2-waypoint path per agent = the (next, next_next) window of PathFinderComponent (Components.h:96-119)
    void uploadWindows() {
        auto& seek = getComponents();
        const uint32_t n = (uint32_t)seek.size();
        std::vector<uint32_t> off(n + 1);
        std::vector<float> wp(4 * (size_t)n), end(2 * (size_t)n), r_next(2 * (size_t)n);
        std::vector<RtsEdge> portals(2 * (size_t)n);
        std::vector<int32_t> path_id(n), waypoint(n, 0);
        for (uint32_t i = 0; i < n; ++i) {
            const auto& c = seek[i];
            off[i] = 2 * i;
            wp[4 * i] = c.r_next.x; wp[4 * i + 1] = c.r_next.y; wp[4 * i + 2] = c.r_next_next.x; wp[4 * i + 3] = c.r_next_next.y;
            portals[2 * i] = {c.portal_next.from.x, c.portal_next.from.y, c.portal_next.t.x, c.portal_next.t.y, c.portal_next.l};
            portals[2 * i + 1] = {c.portal_next_next.from.x, c.portal_next_next.from.y, c.portal_next_next.t.x, c.portal_next_next.t.y, c.portal_next_next.l};
            end[2 * i] = c.path_end.x; end[2 * i + 1] = c.path_end.y;
            r_next[2 * i] = c.r_next.x; r_next[2 * i + 1] = c.r_next.y;
            path_id[i] = c.state == MoveState::MOVING ? (int32_t)i : -1;
        }
        off[n] = 2 * n;
        RtsPathTable t{n, off.data(), wp.data(), portals.data(), end.data(), nullptr, 0};
        std::vector<int32_t> grp;
        if (command_group.size() == n) {   // groups -1 (no command) get ids of their own behind the real ones
            int32_t ng = 0;
            for (auto g : command_group) ng = std::max(ng, g + 1);
            grp.resize(n);
            for (uint32_t i = 0; i < n; ++i) grp[i] = command_group[i] >= 0 ? command_group[i] : ng++;
            t.group = grp.data();
            t.n_groups = (uint32_t)ng;
        }
        check(rts_upload_paths(h, &t), "rts_upload_paths");
        RtsAgentView v{};
        v.path_id = path_id.data(); v.waypoint = waypoint.data(); v.r_next = r_next.data();
        check(rts_update_agents(h, &v), "rts_update_agents");
        paths_dirty = false;
    }
};
