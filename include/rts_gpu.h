/*
 * rts_gpu.h — C ABI of the B200-native agent-update hot path of RTSEngine.
 *
 * One call to rts_step() advances every agent by one 1/60 s frame and replaces, for the
 * Physics and Seek systems, what ECSystem::update does on the CPU
 * (reference: src/ECS.cpp:65-113):
 *
 *   bin agents into the neighbour-search grid        src/Utils/NeighbourSearcherContext.cpp:18-64
 *   physics neighbour pass  (|dr|^2 < r_max2)        src/Utils/NeighbourSearcherStrategy.cpp:10-68
 *   boids forces, clamp, inertia decay               src/Systems/PhysicsSystem.cpp:16-117
 *   wall constraint, applied twice                   src/Systems/PhysicsSystem.cpp:130-162, src/Edges.cpp:245-296
 *   seek toward the funnel waypoint, turn, advance   src/Systems/SeekSystem.cpp:76-111,181-233,277-304,344-385
 *   merge velocities, integrate                      src/Systems/{Physics,Seek}System.cpp communicate(), src/ECS.cpp:92-113
 *   point-in-triangle walk on the CDT                src/PathFinding/Triangulation.cpp:1497-1567
 *
 * The structs are plain data; every pointer is a HOST pointer unless its name ends in _dev.
 * A handle is single-caller (the reference drives its systems from one thread, src/main.cpp:80).
 * All functions return RTS_OK or a negative RtsStatus; rts_last_error() gives the text.
 * There is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef RTS_GPU_H
#define RTS_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RtsStatus {
    RTS_OK = 0,
    RTS_E_CUDA = -1,     /* a CUDA runtime call failed                                   */
    RTS_E_NCCL = -2,     /* an NCCL call failed                                          */
    RTS_E_ARG = -3,      /* bad argument (null pointer, size mismatch, out of range)     */
    RTS_E_CAPACITY = -4, /* a fixed-capacity buffer overflowed (halo / migration)        */
    RTS_E_STATE = -5     /* call order violated (step before map/agents are uploaded)    */
} RtsStatus;

/* MoveState, src/core.h:311-315 */
enum { RTS_MOVING = 0, RTS_STANDING = 1, RTS_HOLDING = 2 };

/* per-agent flag bits reported by rts_download (RtsAgentView.flags) */
enum {
    RTS_FLAG_REPLAN = 1u,     /* updateTarget() shifted the window and asked for a re-plan (SeekSystem.cpp:294-303) */
    RTS_FLAG_WALK_FAILED = 2u /* triangle walk hit its hop limit or an inconsistent table                           */
};

/* Simulation parameters. rts_default_params() fills the reference's values. */
typedef struct RtsParams {
    /* MapGrid (wall lookup grid): cells of map_cell_size units, src/Geometry.h:3-7 */
    int32_t map_nx, map_ny;
    float map_cell_size; /* 5 */
    /* physics neighbour-search grid: ns_n = int(W / ns_cell_size) + 1, cell = 20*RHARD = 60
       (NeighbourSearcherContext.cpp:70-71, PhysicsSystem.cpp:5-9) */
    int32_t ns_nx, ns_ny;
    float ns_cell_size;
    float r_max2; /* 30: "10*RHARD" is passed where a squared radius is expected (Context.cpp:13) */
    /* triangle cache grid (SearchGrid of the Triangulation), src/main.cpp:50 */
    int32_t tri_nx, tri_ny;
    float tri_cell_size; /* 40 */
    /* PhysicsSystem::force_multipliers, defaults from src/UI.cpp:124-131 */
    float mul_push, mul_repulse, mul_scatter, mul_align, mul_seek, mul_decay, mul_velocity;
    float dt;           /* 1/60, src/ECS.h:18 */
    float rhard;        /* 3,  src/core.h:95 */
    float rscatter;     /* 50, src/core.h:94 */
    float sys_max_speed;/* PhysicsSystem::max_speed = 25 (PhysicsSystem.h:37); inert clamp at PhysicsSystem.cpp:27-30 */
    /* SeekSystemSettings, src/Settings.cpp:53-56 (used by the arrival rule) */
    float finish_angle, finish_radius, finish_n_steps, finish_cone_len;
    /* behaviour switches */
    uint32_t enable_arrival; /* 0: SeekSystem::finalizeSeek is skipped; 1: arrival rule with frame-start snapshot semantics:
                                a MOVING agent on its last leg within 2*radius + sqrt(finished_area(group)/pi) of path_end stops
                                and stops its same-group neighbours of the seek neighbour list within finish_radius
                                (SeekSystem.cpp:235-275, CommandGroupManager.h:141-177); the reference runs this inside its serial
                                per-agent loop, here every agent sees the group areas and neighbour states of the frame start.
                                On slab handles the halo band widens to finish_radius (< ns_cell_size required), halo records
                                carry the navigation record so that every rank evaluates the rule for its ghosts too and stops
                                only agents it owns, and the per-group stop COUNTS are all-reduced every frame (ncclAllReduce,
                                the only collective on the frame path; such frames are not replayed from a CUDA graph): the
                                result stays bit-identical to the single-domain run for any number of slabs */
    uint32_t walk_all;       /* 1: keep tri_idx for every agent; 0: MOVING agents only (as SeekSystem.cpp:206) */
    /* multi-GPU slab owned by this handle: physics-grid rows [slab_row_lo, slab_row_hi); 0,ns_ny = whole map */
    int32_t slab_row_lo, slab_row_hi;
    /* tuning words, 0 = default: [0] 1 = do not keep the per-step merged velocity (RtsAgentView.vel reads 0)
       [1] fine gather-cell size as float bits (default 6.0, clamped to >= 1.05*sqrt(r_max2))
       [2] 1 = launch kernels one by one instead of replaying a CUDA graph of eight frames
       [3] unused (round 1: a frame-kernel variant that staged each CTA's neighbour window in shared memory; it measured
           20-35 % slower than the L1-served path and was removed)
       [4] 1 = search the walk queue with a k_walk launch right after k_scatter instead of leaving it to the leading CTAs of
           the next frame kernel
       [5] 8 = slab handles step their edge rows first on a comm stream, followed there by the exchange and the ingest,
           while the interior is stepped on the main stream (measured equal to the plain layout);
           16 = slab frames are replayed from a CUDA graph as well, ncclSend/ncclRecv captured with them
           32 = keep the ncclSend/ncclRecv exchange instead of the peer-memory exchange (rts_slab_exchange_mode)
           64 = launch the frame kernels without programmatic dependent launch (by default every kernel of the frame chain
                is launched with cudaLaunchAttributeProgrammaticStreamSerialization and waits in griddepcontrol.wait) */
    uint32_t reserved[6];
    /* arithmetic of the frame kernel:
       RTS_PRECISION_EXACT (0, the default): IEEE division / square root, no FMA contraction, neighbour sums added in the
         reference's visit order — every output is bit-identical to an IEEE build of the reference (the parity configuration);
       RTS_PRECISION_FAST (1): tolerance mode — same neighbour sets, same wall contacts, same waypoint / state / angle
         decisions and cell / triangle indices that are exact functions of the positions it produces, but the force terms
         use rsqrt.approx / rcp.approx, fused multiply-adds and are added in memory order: after one frame from the same
         state positions agree to ~1e-7 relative and velocities to 1e-5 of max(|v|, 1) except where an agent's force terms
         cancel each other (error bounded by 2e-6 of the sum of the term magnitudes). Results are not
         reproducible bit for bit from run to run (the memory order within a fine cell comes from atomics). */
    uint32_t precision;
    /* Alignment term (Reynolds velocity matching) — an EXTENSION, off by default and in every parity run: the reference
       carries the ALIGN multiplier (UI.cpp:124), the radius RALLIGN (core.h:97) and the per-neighbour velocity difference
       InteractionData::dv (ECS.h:473-483), but the code that would use them is commented out (PhysicsSystem.cpp:87-91).
       With enable_alignment = 1, for a MOVING agent i:  inertia_vel += mul_align * mean(v_j - v_i)  over the physics
       neighbours j that are MOVING and satisfy x > r_collision^2 / ralign^2 (the predicate of the commented code), v = the
       merged velocity of the previous frame. Single-domain handles that keep the per-step velocity only. */
    uint32_t enable_alignment;
    float ralign; /* 50, src/core.h:97 */
} RtsParams;
enum { RTS_PRECISION_EXACT = 0, RTS_PRECISION_FAST = 1 };

/* CDT triangle, counter-clockwise; neighbours[k] is across edge (verts[k], verts[k+1]).
   Mirrors `Triangle` of src/PathFinding/Triangulation.h:35-48 (44 B there, 48 B here). */
typedef struct RtsTriangle {
    int32_t vx[3], vy[3];
    uint32_t neighbours[3]; /* 0xFFFFFFFF = none */
    uint8_t is_constrained[3];
    uint8_t type;
    uint32_t pad_[2];
} RtsTriangle;

/* wall segment, `Edgef` of src/core.h:217-232; outward normal is (t.y, -t.x) */
typedef struct RtsEdge {
    float from_x, from_y, t_x, t_y, l;
} RtsEdge;

/* The four orientation-indexed 1-D edge finders of `Edges` (src/Edges.h:15-46, EdgesFinder1D.h:11-48)
   flattened to CSR: line_offsets[o] has n_lines[o]+1 entries into edges[o]; every line keeps the
   order of edge_array2_[line] (sorted by the map builder, MapGrid.cpp:1277-1303). */
typedef struct RtsWallTable {
    const RtsEdge* edges[4];
    const uint32_t* line_offsets[4];
    uint32_t n_lines[4]; /* ny+1, nx+ny+1, nx+1, nx+ny+1  (Edges.cpp:6-19) */
} RtsWallTable;

/* Funnel paths. Path p owns waypoints [offsets[p], offsets[p+1]); waypoint k carries the portal
   the agent crosses to reach it (PathFinder2::PathAndPortals, PathFinder2.h:62-66). An agent at
   waypoint index w steers to r_next (initially waypoint w) with portal_next = portal w,
   r_next_next = waypoint min(w+1, last), portal_next_next likewise: exactly the two-entry window
   of PathFinderComponent (src/Systems/Components.h:96-119).  A drop-in host that keeps the
   reference planner uploads one 2-waypoint path per agent. */
typedef struct RtsPathTable {
    uint32_t n_paths;
    const uint32_t* offsets;  /* n_paths + 1 */
    const float* waypoints;   /* 2 floats per waypoint */
    const RtsEdge* portals;   /* one per waypoint */
    const float* path_end;    /* 2 floats per path: the commanded destination (PathFinderComponent::path_end) */
    const int32_t* group;     /* command group of each path (CommandGroupManager2::entity2command_group); NULL: group = path */
    uint32_t n_groups;        /* number of command groups (ignored when group is NULL) */
} RtsPathTable;

/* per-unit-type constants (UnitTypes.txt + component defaults, Components.h:84-119) */
typedef struct RtsUnitType {
    float lambda;         /* 0.169 inertia decay        */
    float max_speed;      /* 25   PhysicsComponent      */
    float seek_max_speed; /* 50   PathFinderComponent   */
    float turn_rate;      /* 5 degrees / frame          */
    float desired_angle;  /* 0                          */
    float pad_[3];
} RtsUnitType;

/* Agent state, structure of arrays in the caller's component order (index = component index of the
   reference's dense component vectors; it defines the summation order of the force pass).
   In rts_set_agents every pointer except r may be NULL (defaults in brackets). In rts_download a
   NULL pointer means "not wanted". */
typedef struct RtsAgentView {
    float* r;            /* 2n  position                                               */
    float* vel;          /* 2n  download only: merged velocity of the last step before ECS.cpp:112 zeroes it */
    float* inertia;      /* 2n  PhysicsComponent::inertia_vel                     [0]  */
    float* angle;        /* n   degrees                                           [0]  */
    float* angle_vel;    /* n                                                     [0]  */
    float* radius;       /* n   >= 0 (the sign bit is ignored)                     [3]  */
    float* mass;         /* n   >= 0 (the sign bit is ignored)                     [1]  */
    uint8_t* state;      /* n   MoveState                                  [STANDING]  */
    uint8_t* player;     /* n                                                     [0]  */
    uint8_t* unit_type;  /* n   index into the unit-type table                    [0]  */
    int32_t* path_id;    /* n   index into the path table, -1 = none             [-1]  */
    int32_t* waypoint;   /* n   waypoint index within the path                    [0]  */
    float* r_next;       /* 2n  current steering target; NaN = take it from the path table [NaN] */
    uint32_t* gid;       /* n   global id: orders neighbours inside a cell (= component index) [index] */
    uint32_t* tri_idx;   /* n   download only: containing CDT triangle                 */
    uint32_t* ns_cell;   /* n   download only: physics neighbour-grid cell (Grid::coordToCell) */
    uint32_t* map_cell;  /* n   download only: MapGrid cell (Grid::coordToCell, 5-unit grid)   */
    uint32_t* flags;     /* n   download only: RTS_FLAG_*                               */
} RtsAgentView;

typedef struct RtsTimings {
    double last_step_ms; /* CUDA-event time of the whole last rts_step call        */
    double kernel_ms[8]; /* totals in profiling mode: [0] step [1] bin [2] scan [3] reorder */
    uint64_t kernel_launches;
    uint64_t steps;
} RtsTimings;

typedef struct RtsContext* RtsHandle;

/* reference defaults for a map of map_nx × map_ny cells (5 units each) */
void rts_default_params(RtsParams* p, int32_t map_nx, int32_t map_ny);

/* replaces: construction of PhysicsSystem / SeekSystem (PhysicsSystem.cpp:5-9, SeekSystem.cpp:5-10) */
int32_t rts_create(const RtsParams* params, int32_t device_id, RtsHandle* out);
int32_t rts_destroy(RtsHandle h);
const char* rts_last_error(RtsHandle h);
/* replaces: the UI writing PhysicsSystem::force_multipliers / SeekSystem::settings_ (UI.cpp:121-165) */
int32_t rts_set_params(RtsHandle h, const RtsParams* params);

/* replaces: the tables PhysicsSystem::p_map_ (Edges*) and SeekSystem::p_cdt_ (Triangulation*) point at;
   call again after Game::updateTriangulation (Game.cpp:38-48). cell2tri has tri_nx*tri_ny entries; NULL = build the
   cache on the device from the triangles (rts_build_cell_grid with last_found_tri = 0). */
int32_t rts_upload_map(RtsHandle h, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                       const RtsWallTable* walls);
/* replaces: Triangulation::updateCellGrid (Triangulation.cpp:1829-1864), the host loop that rebuilds cell2triangle_ind_
   after every CDT change with one findTriangle(centre, from_last_found=true) per cache cell in zig-zag order. Rebuilds the
   handle's cache from the uploaded triangles, bit-identical to the reference's table; last_found_tri is
   Triangulation::last_found_tri_ind_ on entry (Triangulation.h:147; it only matters when the first cell centre lies on a
   triangle edge). cell2tri_out (host, tri_nx*tri_ny entries) may be NULL. */
int32_t rts_build_cell_grid(RtsHandle h, uint32_t last_found_tri, uint32_t* cell2tri_out);
/* replaces: the SAME hook after ONE building insert (Game.cpp:38-48 -> MapGrid.cpp:1095-1312): the CDT changes in a small
   neighbourhood and eight walls join eight lines of the edge finders; only that travels instead of the whole table set.
     n_tris_total                      size of Triangulation::triangles_ after the insert (>= before: inserts only append)
     tri_indices / tri_records [n_changed]   every triangle whose record differs from what the handle holds, incl. all new ones
     line_orient / line_index / line_counts [n_lines_changed], line_edges   the COMPLETE new content of every wall line that
                                       changed (edge_array2_[line] after the re-sort, MapGrid.cpp:1277-1303), concatenated
     last_found_tri                    Triangulation::last_found_tri_ind_ (only matters if cache cell 0 is affected)
   The handle patches the triangle table, appends the changed lines to its wall pool, rebuilds the scan lists of the cache
   cells the old and new records touch and re-runs Triangulation::updateCellGrid for that band of cache rows — the
   cell -> triangle cache stays identical to the reference's full rebuild. RTS_E_CAPACITY: the slack of the device tables
   is used up, call rts_upload_map. */
int32_t rts_update_map_region(RtsHandle h, uint32_t n_tris_total, const uint32_t* tri_indices, const RtsTriangle* tri_records,
                              uint32_t n_changed, const uint32_t* line_orient, const uint32_t* line_index, const uint32_t* line_counts,
                              const RtsEdge* line_edges, uint32_t n_lines_changed, uint32_t last_found_tri);
/* host-to-device bytes of the last rts_upload_map / rts_update_map_region call, and of the last full rts_upload_map */
int32_t rts_map_upload_bytes(RtsHandle h, uint64_t* last_call, uint64_t* last_full_upload);
/* the handle's current cell -> triangle cache (tri_nx * tri_ny entries), without rebuilding it */
int32_t rts_get_cell_grid(RtsHandle h, uint32_t* cell2tri_out);
int32_t rts_upload_unit_types(RtsHandle h, const RtsUnitType* types, uint32_t n_types);
/* replaces: PathFinder2::setPathOfAgent writing windows into PathFinderComponents (PathFinder2.cpp:326-380) */
int32_t rts_upload_paths(RtsHandle h, const RtsPathTable* paths);

/* replaces: System2::addComponent for every agent (ECS.h:165-174) — sets the whole population */
int32_t rts_set_agents(RtsHandle h, uint32_t n, const RtsAgentView* agents);
/* replaces: System2::addComponent — appends m agents at indices n..n+m-1 */
int32_t rts_append_agents(RtsHandle h, uint32_t m, const RtsAgentView* agents);
/* replaces: System2::remove (swap-with-last, ECS.h:197-214); indices are component indices */
int32_t rts_remove_agents(RtsHandle h, const uint32_t* indices, uint32_t m);
/* replaces: ECSystem::setMoveState + SeekSystem::issuePaths2 (Game.cpp:211-231): assign path/state to m agents */
int32_t rts_command_agents(RtsHandle h, const uint32_t* indices, uint32_t m, int32_t path_id, uint8_t state);
/* replaces: System2::updateSharedData (PhysicsSystem.cpp:165-173, SeekSystem.cpp:307-316): overwrite the non-NULL
   fields of every agent from caller-order arrays (the ECS blackboard), then re-bin */
int32_t rts_update_agents(RtsHandle h, const RtsAgentView* fields);
/* the same for the m agents the host actually touched since the last frame (a move command, an edit): `indices` are
   distinct component indices, every non-NULL array of `fields` holds m entries in the order of `indices`. Only those
   m records cross the bus; the population is re-binned only if positions are among the fields. This is what a drop-in
   System2::updateSharedData sends when it tracks which entities the game changed instead of copying the whole blackboard
   (PhysicsSystem.cpp:165-173 copies all of it every frame because on the CPU that is free). */
int32_t rts_update_agents_indexed(RtsHandle h, const uint32_t* indices, uint32_t m, const RtsAgentView* fields);
uint32_t rts_agent_count(RtsHandle h);
/* CommandGroupManager2::group2finished_surface_area_ (CommandGroupManager.h:133): read / overwrite, n = number of groups */
int32_t rts_get_finished_area(RtsHandle h, float* out, uint32_t n);
int32_t rts_set_finished_area(RtsHandle h, const float* area, uint32_t n);

/* replaces: PhysicsSystem::update + SeekSystem::update + communicate + avoidWall + integration
   (ECS.cpp:65-113), n_steps frames, asynchronous on the handle's stream */
int32_t rts_step(RtsHandle h, uint32_t n_steps);
/* (The point-in-triangle search of the last frame's few unconfirmed agents is finished by the next rts_step — inside its
   frame kernel — or by whichever call reads or replaces the indices first: rts_sync, an rts_download that asks for tri_idx
   or flags, the uploads that edit positions. A download of other fields only — the per-frame communicate of the adapter —
   leaves it to the next frame.) */
/* replaces: System2::communicate → SharedData (ECS.h:136-145); synchronises.  Arrays hold rts_agent_count() entries
   (for a slab handle: call rts_sync first, the count then includes ghost slots, which come back with gid 0xFFFFFFFF) */
int32_t rts_download(RtsHandle h, const RtsAgentView* out);
/* One entry of rts_download_delta: component `index` now has these values of the slow-changing SharedData fields. */
typedef struct RtsAgentDelta {
    uint32_t index;
    float angle_vel;          /* SharedData::angle_vel (SeekSystem.cpp:322-342) */
    float target_x, target_y; /* SharedData::target = r_next                    */
    uint32_t state;           /* MoveState                                      */
} RtsAgentDelta;
/* replaces: the same System2::communicate (ECS.h:233), as a DIFFERENCE against what this handle's previous
   rts_download_delta reported — the mirror image of rts_update_agents_indexed. The frame's merged, wall-projected velocity
   changes for every agent and travels in full (vel[n][2]); angle_vel, state and target change for a few per cent of the agents
   per frame (a turn finished, a waypoint reached, a stop), and only those entities travel: one RtsAgentDelta for every agent
   whose value of any of the three differs BITWISE from the value last reported for it. Applying the entries to the blackboard
   of the previous frame gives exactly the blackboard a full rts_download would give (tests/test_gpu_parity.py). The first call
   after the population changed (rts_set_agents / append / remove) reports every agent.
     changes [cap], n_changes   RTS_E_CAPACITY if more than cap agents changed (the next call then reports everyone again):
                                cap = rts_agent_count() always suffices
     angle_vel, state, target   optional blackboard arrays in component order ([n], [n] bytes, [n][2]): when given, the library
                                writes the entries into them itself (rts_apply_delta) while the velocities are still arriving
   Single-domain handles with component-index gids. Leaves a pending walk queue to the next frame (no index is read).
   Synchronises. */
int32_t rts_download_delta(RtsHandle h, float* vel, RtsAgentDelta* changes, uint32_t cap, uint32_t* n_changes,
                           float* angle_vel, uint8_t* state, float* target);
/* Host-side helper (no device work): writes the entries of a change list into blackboard arrays in component order —
   angle_vel[n], state[n] (MoveState as a byte), target[n][2]; any of the three may be NULL. */
int32_t rts_apply_delta(const RtsAgentDelta* changes, uint32_t n_changes, uint32_t n, float* angle_vel, uint8_t* state, float* target);
int32_t rts_sync(RtsHandle h);
int32_t rts_get_timings(RtsHandle h, RtsTimings* out);
/* device time (ms, CUDA events on the handle's stream) of the last `count` kernels of kind `which`
   0 = fused step kernel, 1 = bin/count, 2 = scan, 3 = reorder; used by bench.py for the roofline line */
int32_t rts_kernel_time_ms(RtsHandle h, int32_t which, double* avg_ms, uint64_t* launches);
int32_t rts_reset_timings(RtsHandle h);
/* on: record a CUDA-event pair around every kernel (and launch kernels one by one, no graph) */
int32_t rts_set_profiling(RtsHandle h, int32_t on);
/* sizeof() of the ABI structs as compiled into the library: 0 RtsParams 1 RtsTriangle 2 RtsEdge 3 RtsWallTable
   4 RtsPathTable 5 RtsUnitType 6 RtsAgentView 7 RtsTimings — lets a binding check its mirror */
size_t rts_abi_sizeof(int32_t which);

/* stand-alone data-parallel queries (each replaces a per-point loop over the reference function) */
/* Grid::coordToCell (Grid.cpp:18-24) for n points on an nx×ny grid of cell size cs */
int32_t rts_coord_to_cell(RtsHandle h, const float* xy, uint32_t n, int32_t nx, int32_t ny, float cs, uint32_t* out);
/* Triangulation::findTriangle(Vector2f,false) (Triangulation.cpp:1588-1590) against the uploaded CDT */
int32_t rts_find_triangles(RtsHandle h, const float* xy, uint32_t n, uint32_t* out);
/* Edges::calcContactEdgesO (Edges.cpp:245-296): per point the number of contact walls, and up to `cap`
   of them in visit order (5 floats each) */
int32_t rts_contact_edges(RtsHandle h, const float* xy, const float* radius, uint32_t n, uint32_t cap,
                          uint32_t* counts, float* edges_out);

/* ---- neighbour queries of the callers of the frame path (first slice), on the cell order the frame keeps ----
   NeighbourSearcherContext::getNeighboursIndsFull (NeighbourSearcherContext.h:44-92; the game's click selection runs it on
   the physics grid, Game.cpp:324) for n query points: per query the number of agents with dist^2 <= r_max^2 and up to `cap`
   of their component indices in the reference's visit order (cells row-major, ascending index inside a cell); a query with
   more hits than `cap` reports the true count and an unspecified selection of `cap` of them. */
int32_t rts_query_radius(RtsHandle h, const float* xy, const float* r_max, uint32_t n, uint32_t cap, uint32_t* counts,
                         uint32_t* indices_out);
/* NeighbourSearcherStrategyAttack::execute (NeighbourSearcherStrategy.cpp:125-177) for every agent, AS WRITTEN: the reference
   never lowers min_enemy_dist_sq from Geometry::BOX[0] (2560), so the "nearest enemy" of an agent is the LAST agent of another
   player with dist^2 < range2 in visit order (calcNearestCells cells, own cell last). out[i] = its component index or -1
   (rts_agent_count() entries). The oracle's version is pinned on the reference's own context (tests/test_oracle_vs_ref.py). */
int32_t rts_attack_targets(RtsHandle h, float range2, int32_t* out);

/* ---- multi-GPU: spatial slab decomposition, one handle (one process, one GPU) per y-slab -----------------------
   The reference has no distributed path; this is the north-star extension.  A handle created with
   params.slab_row_lo/hi owns the physics-grid rows [lo, hi).  Every frame the frame kernel packs (a) agents whose
   new position left the slab (full record, "migrants") and (b) agents within the interaction range of a slab
   boundary (neighbour record only, "halo") into fixed-capacity device buffers; neighbours exchange them with
   ncclSend/ncclRecv over NVLink on the handle's stream; arrivals are appended before the counting sort.
   Agents carry global ids (RtsAgentView.gid, required) so that the summation order — hence every bit of the
   result — is independent of the number of slabs. */
int32_t rts_slab_configure(RtsHandle h, int32_t rank, int32_t nranks, uint32_t cap_migrants, uint32_t cap_halo);
/* NCCL bootstrap: rank 0 creates the 128-byte ncclUniqueId, the host distributes it (torch.distributed, MPI, a file),
   every rank calls rts_comm_init.  libnccl.so.2 is resolved with dlopen at this point, not at load time. */
int32_t rts_comm_unique_id(void* out128);
int32_t rts_comm_init(RtsHandle h, const void* id128);
/* how this handle exchanges with its neighbours: 0 = not a slab handle, 1 = ncclSend/ncclRecv of the exchange buffers every
   frame, 2 = peer exchange — rts_comm_init also maps the neighbours' receive buffers into this process (CUDA IPC handles
   gathered with ncclAllGather) and a small kernel then writes this rank's halo / migrant records straight into them over
   NVLink, exactly as many as were packed, with an epoch flag per direction; the frame has no NCCL call left in it and is
   replayed from a CUDA graph like the single-GPU frame. Falls back to 1 if any rank cannot map its neighbour (or with
   RtsParams.reserved[5] bit 5 (value 32) set). Collective: every rank must call rts_comm_init. */
int32_t rts_slab_exchange_mode(RtsHandle h);
/* after rts_set_agents on every rank: exchange the halo bands once so the first frame sees its ghosts */
int32_t rts_halo_refresh(RtsHandle h);
/* split-phase variants for a host that moves the buffers itself (and for emulating two ranks on one device):
   rts_step_begin | rts_halo_begin  ->  move send buffers to the neighbours' receive buffers  ->  rts_step_end */
int32_t rts_step_begin(RtsHandle h);
int32_t rts_halo_begin(RtsHandle h);
int32_t rts_step_end(RtsHandle h);
int32_t rts_slab_exchange_local(RtsHandle lower, RtsHandle upper);

/* Slab rebalancing (boundaries follow agent-count quantiles, re-evaluated every K frames — crowds that drain
   through corridors pile up in a few slabs otherwise). rts_slab_row_counts: owned agents per physics-grid row (ns_ny
   entries); the host sums them over the ranks and picks new boundaries. rts_slab_set_rows: this handle now owns rows
   [row_lo, row_hi) — call it on both sides of a moved boundary between the same two frames. The agents of the rows that
   changed hands migrate during the next frame through the ordinary migrant buffers (so at most cap_migrants of them per
   direction; the host moves a boundary by fewer rows if they would not fit). Results stay bit-identical to the
   single-domain run: ownership never enters the arithmetic. */
int32_t rts_slab_row_counts(RtsHandle h, uint32_t* counts);
int32_t rts_slab_set_rows(RtsHandle h, int32_t row_lo, int32_t row_hi);

/* raw CUDA stream of the handle (cudaStream_t as void*), so a host that owns device buffers can order work after it */
void* rts_stream(RtsHandle h);

#ifdef __cplusplus
}
#endif
#endif /* RTS_GPU_H */
