// Kernels of the agent-update hot path (sm_100a). See DESIGN.md for the data layout and the
// per-kernel byte accounting.
//
//   k_bin      : fine-cell id + slot rank of freshly uploaded agents (warp-aggregated atomics)
//   k_scan_*   : exclusive scan of the per-cell counts -> cell_start
//   k_scatter  : counting-sort scatter of the neighbour record {pos4,key2} into cell order, builds src[]
//   k_walk     : findTriangle for the agents k_scatter queued (normally done by the leading CTAs of the next k_step)
//   k_step     : the fused frame — neighbour gather, boids forces, clamp/decay, wall pass, seek, merge,
//                wall pass, integrate, slab packing, re-bin of the new position
//   k_step_dense : the same for the crowded agents k_step queues (one warp per agent)
//   k_arrive / k_area_sync / k_area_sync_counts : arrival rule (finalizeSeek) with frame-start snapshot semantics, also across slabs
//   k_ingest_remote / k_halo_seed : slab exchange (the frame's counts published, arrivals appended; halo bands packed without a frame)
//   k_cellgrid_classify / k_cellgrid_resolve : Triangulation::updateCellGrid (cell -> triangle cache)
//   k_border_cache : findTriangle's answers on the far border lines of the cache grid, recorded after every map change
//   k_gather   : slot order -> caller order for rts_download
//   k_gather_delta : ... as a difference against what the host was last told (rts_download_delta)
//   k_apply    : caller order -> slot order for rts_update_agents
//   k_coord_to_cell / k_find_triangles / k_contact_edges : stand-alone point queries
#pragma once

#include "rts_device.cuh"

namespace rts {

#ifndef RTS_USE_DENSE_KERNEL
#define RTS_USE_DENSE_KERNEL 1
#endif
constexpr bool USE_DENSE_KERNEL = RTS_USE_DENSE_KERNEL != 0;
#ifndef RTS_STEP_BLOCK
#define RTS_STEP_BLOCK 64
#endif
constexpr int STEP_BLOCK = RTS_STEP_BLOCK;   // threads per CTA of k_step. 64 x 18 CTAs per SM (55 registers, 8 KB of shared memory each) measured
                                            // 9 % faster over frames 20-220 of config 3 than 128 x 8 with a 16-entry list and a spill list
#ifndef RTS_STEP_MIN_BLOCKS
#define RTS_STEP_MIN_BLOCKS 18
#endif
constexpr int STEP_MIN_BLOCKS = RTS_STEP_MIN_BLOCKS;
#ifndef RTS_NBR_CAP
#define RTS_NBR_CAP 16
#endif
constexpr int NBR_CAP = RTS_NBR_CAP;   // neighbours kept sorted in shared memory per agent
#ifndef RTS_FAST_BLOCK
#define RTS_FAST_BLOCK 128
#endif
#ifndef RTS_FAST_MIN_BLOCKS
#define RTS_FAST_MIN_BLOCKS 10
#endif
#ifndef RTS_PERSISTENT_STEP
#define RTS_PERSISTENT_STEP 0
#endif
// k_step as one residency of the device whose warps take 32-slot tiles from a device-wide counter (see k_step). Built to
// remove the idle tail of the 5.7-wave grid (ncu: the average SM is idle for 16 % of the launch); measured SLOWER on config 3
// (k_step 97 us against 74.6 us — also with the counter's round trip really hidden, see the loop head; static striding without
// the counter 91 us; whole CTAs taking 128 slots at a time 80.5 us — profiles/README.md), so the plain grid stays the default
// and this is a build switch for further experiments.
constexpr bool USE_PERSISTENT_STEP = RTS_PERSISTENT_STEP != 0;
constexpr int FAST_BLOCK = RTS_FAST_BLOCK;             // threads per CTA of the tolerance-mode k_step (no shared memory: registers set the occupancy)
constexpr int FAST_MIN_BLOCKS = RTS_FAST_MIN_BLOCKS;
#ifndef RTS_FAST_DENSE_CAND
#define RTS_FAST_DENSE_CAND 256
#endif
constexpr uint32_t FAST_DENSE_CAND = RTS_FAST_DENSE_CAND;   // tolerance mode: agents with more CANDIDATES than this (slots in their 3x3 fine
                                                            // cells; ~12 at config-3 density, ~52 in the corridors of config 5) are
                                                            // handed to k_step_dense_fast, where FAST_DENSE_LANES lanes share one agent's
                                                            // candidates. One thread per agent is the efficient form — neighbouring
                                                            // slots are equally crowded, so warps stay converged — and the hand-over
                                                            // only has to bound the longest single thread (measured: 64 -> 256 takes
                                                            // config 5 from 0.85 to 0.69 ms per frame and config 3's late frames from
                                                            // 0.122 to 0.107)
#ifndef RTS_FAST_DENSE_LANES
#define RTS_FAST_DENSE_LANES 16
#endif
constexpr uint32_t FAST_DENSE_LANES = RTS_FAST_DENSE_LANES;
#ifndef RTS_DENSE_MIN
#define RTS_DENSE_MIN 16
#endif
#ifndef RTS_DENSE_GRID_PER_SM
#define RTS_DENSE_GRID_PER_SM 8
#endif
constexpr int DENSE_MIN = RTS_DENSE_MIN;   // agents with more neighbours than this are handed to k_step_dense (one warp each)
constexpr int DENSE_MAX = 256;             // ... up to this many; beyond, k_step itself re-scans (the reference gives up at 500,
                                           // Strategy.cpp:62-65). Two agents per warp share k_step_dense's 16 KB of list space.
constexpr int SPILL_CAP = 104;    // further neighbours parked unsorted in local memory before re-scanning is needed
// (with k_step_dense taking every agent beyond the shared-memory list, the spill list has no user and is compiled out)
#ifdef RTS_FORCE_SPILL
constexpr bool USE_SPILL = true;
#else
constexpr bool USE_SPILL = !(USE_DENSE_KERNEL && RTS_NBR_CAP >= RTS_DENSE_MIN);
#endif
constexpr uint32_t INVALID_CELL = 0xFFFFFFFFu;

// Programmatic dependent launch (see launch_chain in rts_abi.cu). pdl_wait: nothing of the previous kernels' output may be
// read, and nothing they read may be overwritten, before this returns. pdl_trigger: the next kernel of the stream may start
// making its CTAs resident (it blocks in its own pdl_wait until this grid has completed). Both are no-ops in a kernel that
// was launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// key2.y = cy << 16 | cx << 3 | ghost << 2 | state, (cx, cy) = 2-D coordinates of the physics-grid cell (13 bits each)
__device__ __forceinline__ uint32_t pack_cs(uint32_t cx, uint32_t cy, uint32_t ghost, uint32_t state) {
    return (cy << 16) | (cx << 3) | (ghost << 2) | (state & 3u);
}
__device__ __forceinline__ int cs_cx(uint32_t y) { return (int)((y >> 3) & 0x1FFFu); }
__device__ __forceinline__ int cs_cy(uint32_t y) { return (int)(y >> 16); }
// Grid::coordToCell on the physics grid, kept as 2-D coordinates (cell index = cx + cy * ns_nx)
__device__ __forceinline__ uint32_t ns_pack(const RtsParams& P, float x, float y, uint32_t ghost, uint32_t state) {
    return pack_cs(axis_cell(x, P.ns_cell_size, P.ns_nx), axis_cell(y, P.ns_cell_size, P.ns_ny), ghost, state);
}
// is the physics-grid cell in key word `ky` one of the (up to) nine cells calcNearestCells gives an agent in cell
// (cxi, cyi)?  (Grid.cpp:103-118: offsets -1..1 on the cell coordinates, cells outside the grid dropped — no wrapping.)
// Always true for two agents inside the map that are within the interaction range; an agent that was pushed over the map
// edge has a cell index wrapped modulo the grid (Grid.cpp:18-24) and is then NOT a neighbour of the agents next to it.
__device__ __forceinline__ bool cell_is_adjacent(uint32_t ky, int cxi, int cyi) {
    const int dxc = cs_cx(ky) - cxi, dyc = cs_cy(ky) - cyi;
    return dxc >= -1 && dxc <= 1 && dyc >= -1 && dyc <= 1;
}
// pos4 = {x, y, radius, mass}; the MoveState is mirrored in the sign bits of radius (bit 0) and mass (bit 1) — both are
// non-negative quantities — so that the neighbour loop gets the neighbour's state with its position, without a second load
__device__ __forceinline__ float4 pack_pos(float x, float y, float radius, float mass, uint32_t state) {
    return make_float4(x, y, __uint_as_float((__float_as_uint(radius) & 0x7FFFFFFFu) | ((state & 1u) << 31)),
                       __uint_as_float((__float_as_uint(mass) & 0x7FFFFFFFu) | ((state & 2u) << 30)));
}
__device__ __forceinline__ int pos_state(float4 o) { return (int)((__float_as_uint(o.z) >> 31) | ((__float_as_uint(o.w) >> 31) << 1)); }
// nav4.w bits: waypoint 0..15 | unit_type 16..23 | player 24..27 | flags 28..31
__device__ __forceinline__ uint32_t pack_misc(uint32_t wp, uint32_t type, uint32_t player, uint32_t flags) {
    return (wp & 0xFFFFu) | ((type & 0xFFu) << 16) | ((player & 0xFu) << 24) | ((flags & 0xFu) << 28);
}

struct FineGrid {
    float inv_cs;       // 1 / fine cell size (cell size >= sqrt(r_max2), so 3x3 fine cells hold every neighbour)
    int32_t nx, ny;     // fine cells per row / rows held by this handle
    int32_t row0;       // first fine row of this handle's slab (0 for a whole map)
    uint32_t n_cells;   // nx * ny
    uint32_t* count;    // [n_cells]   agents per fine cell for the NEXT order (atomics), zeroed by the scan
    uint32_t* start;    // [n_cells+1] first slot of each fine cell in the CURRENT order
    uint32_t* partial;  // scan scratch (two-kernel scan)
    unsigned long long* status;   // one-pass scan: per tile {flag<<32 | value}
    uint32_t* ticket;             // [0] next tile to hand out, [1] epoch of the flags (bumped by the last tile)
};

__device__ __forceinline__ void fine_coords(const FineGrid& F, float x, float y, int& fx, int& fy) {
    fx = min(max((int)(x * F.inv_cs), 0), F.nx - 1);
    fy = min(max((int)(y * F.inv_cs) - F.row0, 0), F.ny - 1);
}

// Slot rank inside a fine cell with one atomic per distinct cell per warp.
// `active` = lanes of the warp that call this (ballot taken before any divergence).
__device__ __forceinline__ uint32_t warp_aggregated_rank(uint32_t* count, uint32_t cell, unsigned active) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(active, cell);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader && cell != 0xFFFFFFFFu) base = atomicAdd(count + cell, (uint32_t)__popc(peers));   // 0xFFFFFFFF: not binned here
    base = __shfl_sync(peers, base, leader);
    return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

// ---- multi-GPU slab decomposition: exchange buffers of one direction (0 = towards lower rows, 1 = higher) ----
struct ExBuf {
    uint32_t* hdr;      // [0] migrants, [1] halo records, [2] overflow flag, [3] unused
    float4* m_pos4;     // migrants: the full per-agent record
    uint2* m_key2;
    float4* m_vel4;
    float4* m_nav4;
    uint32_t* m_tri;
    float2* m_velout;
    float4* h_pos4;     // halo: the neighbour record ...
    uint2* h_key2;
    float4* h_nav4;     // ... and the navigation record {r_next, path_id, misc}: the arrival rule evaluates ghosts too (k_arrive)
};
// one exchange buffer = [hdr 16 B][m_pos4][m_vel4][m_nav4][h_pos4][h_nav4][m_key2][h_key2][m_velout][m_tri]
__host__ __device__ inline ExBuf carve_exbuf(void* base, uint32_t cap_mig, uint32_t cap_halo) {
    ExBuf B;
    char* p = static_cast<char*>(base);
    B.hdr = reinterpret_cast<uint32_t*>(p); p += 16;
    B.m_pos4 = reinterpret_cast<float4*>(p); p += (size_t)cap_mig * 16;
    B.m_vel4 = reinterpret_cast<float4*>(p); p += (size_t)cap_mig * 16;
    B.m_nav4 = reinterpret_cast<float4*>(p); p += (size_t)cap_mig * 16;
    B.h_pos4 = reinterpret_cast<float4*>(p); p += (size_t)cap_halo * 16;
    B.h_nav4 = reinterpret_cast<float4*>(p); p += (size_t)cap_halo * 16;
    B.m_key2 = reinterpret_cast<uint2*>(p); p += (size_t)cap_mig * 8;
    B.h_key2 = reinterpret_cast<uint2*>(p); p += (size_t)cap_halo * 8;
    B.m_velout = reinterpret_cast<float2*>(p); p += (size_t)cap_mig * 8;
    B.m_tri = reinterpret_cast<uint32_t*>(p);
    return B;
}
__host__ __device__ inline size_t exbuf_bytes(uint32_t cap_mig, uint32_t cap_halo) {
    return (16 + (size_t)cap_mig * (16 + 16 + 16 + 8 + 8 + 4) + (size_t)cap_halo * (16 + 16 + 8) + 15) & ~(size_t)15;
}
struct SlabDev {
    int32_t row_lo, row_hi;   // physics-grid rows owned: [row_lo, row_hi)
    float y_lo, y_hi;         // = row * ns_cell_size
    float halo;               // agents this close to a boundary are mirrored on the neighbour (>= sqrt(r_max2))
    int32_t has[2];           // neighbour present in direction 0 / 1
    uint32_t cap_mig, cap_halo;
    ExBuf send[2], recv[2];
    uint32_t* overflow;       // sticky: 1 = a migrant / halo record did not fit its buffer, 2 = an agent outside the edge rows packed,
                              // 3 = the neighbour's records of this frame never arrived (peer exchange timed out)
    // ---- peer exchange (NVLink, CUDA IPC): the neighbour's receive buffers are mapped into this process and the kernels write
    // records straight into them; a flag per direction carries the epoch (frame counter) of the last complete
    // delivery. Receive buffers are double-buffered by epoch parity: the neighbour may already deliver frame e+1 while
    // frame e is being ingested (it cannot be two frames ahead: its frame e+2 needs this rank's frame e+1 delivery).
    // k_step / k_step_dense* / k_halo_seed write their records directly into the neighbour's buffer (pack_target); the slot
    // counters stay local, and the first thread of k_ingest_remote — the next kernel of the stream — copies them across and
    // raises the flag (publish_to_peers).
    int32_t peer_mode;
    char* recv_base[2][2];    // [direction][parity] this rank's receive buffers (same layout as send / recv: carve_exbuf)
    char* peer_base[2][2];    // [direction][parity] where records sent in `direction` land: the neighbour's recv_base[1 - direction]
    unsigned long long* flag_local[2];  // [direction] peer_word of the last delivery that arrived from that direction
    unsigned long long* flag_peer[2];   // [direction] the neighbour's flag_local[1 - direction]
    uint32_t* epoch;          // [0] exchanges completed (k_scatter ticks it)
    int32_t publish_here;     // (unused since the first thread of k_ingest_remote publishes: publish_to_peers)
    // Edge rows (fine-grid rows relative to FineGrid::row0): slots of rows < edge_row_lo or >= edge_row_hi can send
    // something to a neighbour this frame (halo band + one frame of travel). k_step<PHASE 1> steps exactly those, so that
    // the exchange can start while k_step<PHASE 2> is still working on the interior.
    int32_t edge_row_lo, edge_row_hi;
    int32_t interior_launch;  // set in the copy handed to PHASE 2 launches: an agent that packs there was missed by PHASE 1
    // The PHASE 1 launch has 2*edge_blocks CTAs: the first edge_blocks cover the slots from 0 upward (the low edge rows),
    // the others the slots from the first high edge row upward. Edge slots beyond that coverage are left to PHASE 2 and
    // reported there if they pack.
    uint32_t edge_blocks;
};

// where the records packed for direction `dir` are written: the local send buffer, or — peer exchange — straight into the
// neighbour's receive buffer of this epoch's parity (the slot COUNTERS stay local: send[dir].hdr; publish_to_peers copies them)
__device__ __forceinline__ ExBuf pack_target(const SlabDev& SL, int dir) {
    if (!SL.peer_mode) return SL.send[dir];
    const uint32_t e = *((volatile uint32_t*)SL.epoch);
    return carve_exbuf(SL.peer_base[dir][e & 1u], SL.cap_mig, SL.cap_halo);
}
// Peer exchange, end of the packing. Every thread that wrote a record into a neighbour's buffer has executed a system-scope
// fence behind it (finish_agent, k_halo_seed), and the kernels that pack have ended before k_ingest_remote starts: when its
// first thread runs, every record of the frame has landed. That thread publishes ONE 64-bit word per neighbour —
// {epoch + 1 : 24 | migrants : 20 | halo records : 20}: flag and header in a single store, no second round trip — before
// anything of that kernel waits for the neighbours' words (so no rank waits for a word whose sender is waiting itself).
__device__ __forceinline__ unsigned long long peer_word(uint32_t epoch1, uint32_t nm, uint32_t nh) {
    return ((unsigned long long)(epoch1 & 0xFFFFFFu) << 40) | ((unsigned long long)(nm & 0xFFFFFu) << 20) | (unsigned long long)(nh & 0xFFFFFu);
}
__device__ __forceinline__ void publish_to_peers(const SlabDev& SL) {   // one thread
    __threadfence_system();
    const uint32_t e = *((volatile uint32_t*)SL.epoch);
    for (int dir = 0; dir < 2; ++dir) {
        if (!SL.has[dir]) continue;
        volatile uint32_t* lh = SL.send[dir].hdr;
        *((volatile unsigned long long*)SL.flag_peer[dir]) = peer_word(e + 1u, min(lh[0], SL.cap_mig), min(lh[1], SL.cap_halo));
    }
}

// ... and already by the first thread of k_step_dense* when the frame kernel queued no crowded agent (qn == 0: nothing more will
// be packed in this frame): the word is then on its way to the neighbour while this kernel and the launch of k_ingest_remote are
// still ahead. k_ingest_remote publishes in any case — the same word.
template <bool MULTI>
__device__ __forceinline__ void publish_early(const SlabDev& SL, uint32_t queued) {
    if (MULTI && SL.peer_mode && queued == 0u && blockIdx.x == 0 && threadIdx.x == 0) publish_to_peers(SL);
}

struct AgentArrays {
    uint32_t* n_dev;  // live slots of the CURRENT order (owned agents + ghosts); written by the scan
    // neighbour record, CURRENT cell order
    float4* pos4;   // {x, y, radius, mass}
    uint2* key2;    // {gid, cy<<16 | cx<<3 | ghost<<2 | state}
    uint32_t* src;  // slot -> index into the previous-order arrays below
    // previous-order arrays (outputs of the last step / the upload), read through src[]
    float4* vel4;   // {inertia.x, inertia.y, angle, angle_vel}
    float4* nav4;   // {r_next.x, r_next.y, path_id, misc}
    uint32_t* tri;
    float2* vel_out;  // merged velocity of the last step (only if keep_vel)
    // outputs of this step, index = slot
    float4* pos4n;
    uint2* key2n;
    float4* vel4n;
    float4* nav4n;
    uint32_t* trin;
    float2* vel_outn;
    uint2* cellrank;  // {next fine cell, rank inside it}
    uint32_t* stop_req;  // arrival: 1 = this slot's agent stops this frame (set by k_arrive, consumed by k_step)
    uint32_t* walk_q;    // new slots of the agents whose triangle k_scatter could not confirm from the hint (k_walk's work list)
    uint32_t* walk_n;    // [parity] their number; k_scatter of the next frame resets it
    uint32_t* dense_q;   // slots of the agents deferred to k_step_dense this frame
    uint32_t* dense_n;   // [0] their number (reset by k_scatter); [4] the same for the queue of a slab's edge-row launch
    uint32_t* work;      // tiles handed out by the persistent k_step beyond each warp's first (reset by k_scatter); a cache line of its own
    uint32_t n_tris_dev_hint;  // != 0 once a map is uploaded (a missing triangle index then means the walk failed)
};

// ------------------------------------------------------------------------------------------------
__global__ void k_bin(uint32_t n, RtsParams P, FineGrid F, AgentArrays A) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned active = __ballot_sync(0xffffffffu, p < n);
    if (p >= n) return;
    const float4 pos = A.pos4n[p];
    uint2 k = A.key2n[p];
    k.y = ns_pack(P, pos.x, pos.y, (k.y >> 2) & 1u, k.y & 3u);
    A.key2n[p] = k;
    int fx, fy;
    fine_coords(F, pos.x, pos.y, fx, fy);
    const uint32_t cell = (uint32_t)fy * (uint32_t)F.nx + (uint32_t)fx;
    const uint32_t rank = warp_aggregated_rank(F.count, cell, active);
    A.cellrank[p] = make_uint2(cell, rank);
}

// ---- scan: 2048 cells per CTA (256 threads x 8) ----
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_reduce_sum(uint32_t v, uint32_t* smem) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = v;
    __syncthreads();
    uint32_t t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0u;
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) smem[0] = t;
    }
    __syncthreads();
    t = smem[0];
    __syncthreads();
    return t;
}

// One-pass exclusive scan of the cell counts (decoupled look-back): tiles are handed out by an atomic ticket, so a tile
// only ever waits for tiles that already started. Every tile publishes its AGGREGATE as soon as it is known and its
// INCLUSIVE PREFIX once its own look-back is done; the look-back (warp 0, 32 predecessors per round) stops at the nearest
// predecessor that already has a prefix, so its length does not grow with the number of tiles (46 M cells on the
// config-4 map: 22 000 tiles). Flags carry an epoch, so nothing is reset between launches (the kernel is replayed from a
// CUDA graph with frozen arguments): status word = {flag : 32 | value : 32}, flag = 2*epoch + 1 (aggregate) / + 2 (prefix).
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(FineGrid F, uint32_t* n_out) {
    static_assert(SCAN_ITEMS == 8, "two uint4 per thread");
    __shared__ uint32_t smem[32];
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile, s_prefix, s_epoch;
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) {
        s_epoch = *((volatile uint32_t*)(F.ticket + 1));
        s_tile = atomicAdd(F.ticket, 1u);
    }
    __syncthreads();
    const uint32_t tile = s_tile, epoch = s_epoch;
    // count[] and start[] are padded to n_pad = a multiple of 8 >= n_cells + 1 (padding counts are 0), so every thread
    // moves two aligned uint4 and start[n_cells] = total falls out of the scan
    const uint32_t n_pad = (F.n_cells + 1u + 7u) & ~7u;
    const uint32_t n_tiles = (n_pad + SCAN_TILE - 1) / SCAN_TILE;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    if (base < n_pad) {
        const uint4 a = *reinterpret_cast<const uint4*>(F.count + base), b = *reinterpret_cast<const uint4*>(F.count + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) v[i] = 0u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
    const uint32_t aggregate = block_reduce_sum(s, smem);
    const uint32_t flag_a = 2u * epoch + 1u, flag_p = 2u * epoch + 2u;   // (epoch starts at 0 and status words at 0: flag 0 never matches)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) atomicExch(F.status + tile, ((unsigned long long)(tile == 0u ? flag_p : flag_a) << 32) | aggregate);
    if (w == 0) {
        uint32_t before = 0;
        for (int64_t j0 = (int64_t)tile - 1; j0 >= 0; j0 -= 32) {
            const int64_t j = j0 - lane;          // lane 0 looks at the nearest predecessor
            uint32_t fl = flag_p, val = 0u;       // (beyond tile 0: an empty prefix)
            if (j >= 0) {
                unsigned long long wv;
                do { wv = *((volatile unsigned long long*)(F.status + j)); fl = (uint32_t)(wv >> 32); } while (fl != flag_a && fl != flag_p);   // plain L2 reads
                val = (uint32_t)wv;
            }
            const unsigned has_prefix = __ballot_sync(0xffffffffu, fl == flag_p);
            const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31;   // nearest predecessor whose inclusive prefix is known
            uint32_t c = (lane <= stop) ? val : 0u;
            for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
            before += __shfl_sync(0xffffffffu, c, 0);
            if (has_prefix) break;
        }
        if (lane == 0) {
            s_prefix = before;
            if (tile != 0u) atomicExch(F.status + tile, ((unsigned long long)flag_p << 32) | (before + aggregate));
        }
    }
    // exclusive scan of the per-thread sums
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0u;
        uint32_t ti = t;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
    }
    __syncthreads();
    uint32_t run = s_prefix + warp_tot[w] + (incl - s);
    if (base < n_pad) {
        uint32_t o[SCAN_ITEMS];
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) { o[i] = run; run += v[i]; }
        *reinterpret_cast<uint4*>(F.start + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(F.start + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        *reinterpret_cast<uint4*>(F.count + base) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(F.count + base + 4) = make_uint4(0u, 0u, 0u, 0u);
        if (n_out && base <= F.n_cells && F.n_cells < base + SCAN_ITEMS) *n_out = o[F.n_cells - base];   // = total
    }
    if (tile == n_tiles - 1 && threadIdx.x == 0) {   // every tile has read the epoch and taken its ticket by now
        F.ticket[0] = 0u;
        F.ticket[1] = epoch + 1u;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partial(FineGrid F) {
    __shared__ uint32_t smem[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < F.n_cells) s += F.count[base + i];
    s = block_reduce_sum(s, smem);
    if (threadIdx.x == 0) F.partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(FineGrid F, uint32_t* n_out) {
    __shared__ uint32_t smem[32];
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    // offset of this tile = sum of the partials of all tiles before it
    uint32_t off = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += SCAN_THREADS) off += F.partial[b];
    off = block_reduce_sum(off, smem);
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < F.n_cells) ? F.count[base + i] : 0u;
        s += v[i];
    }
    // exclusive scan of the per-thread sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0u;
        uint32_t ti = t;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;
    }
    __syncthreads();
    uint32_t run = off + warp_tot[w] + (incl - s);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < F.n_cells) {
            F.start[base + i] = run;
            F.count[base + i] = 0u;
            run += v[i];
        }
    }
    if (base <= F.n_cells - 1 && F.n_cells - 1 < base + SCAN_ITEMS) {  // owner of the last cell closes the CSR
        F.start[F.n_cells] = run;
        if (n_out) *n_out = run;
    }
}

// Counting-sort scatter of the neighbour record into the new cell order; builds src[].
// ... and the first half of the triangle walk: the triangle of the previous frame (carried forward by k_step into
// trin[p]) is a hint. A position STRICTLY inside a triangle is contained in no other one, so findTriangle — whatever cache
// cell it starts from, however it walks, or through its linear scan — can only end there: the entry is already right.
// Agents move < 2 units per frame, so nearly all of them pass; the others (hint missing, on an edge, moved on) are queued
// for k_walk, which runs the reference's own search for them.
__global__ void __launch_bounds__(256) k_scatter(uint32_t n, FineGrid F, AgentArrays A, SlabDev SL, int multi, RtsParams P, MapDev M,
                                                 int parity, int do_walk, int tick_epoch) {
    pdl_wait();
    pdl_trigger();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0 && tick_epoch) SL.epoch[0] += 1u;   // one peer exchange has been delivered and ingested (see SlabDev)
    if (p == 0) {
        A.dense_n[0] = 0u; A.dense_n[4] = 0u;   // the dense queues of this frame (whole / interior, edge rows) have been consumed
        A.work[0] = 0u;                         // ... and the tile counter of its frame kernel
        A.walk_n[parity ^ 1] = 0u;              // ... and the walk queue of the previous frame (its k_walk was joined before this launch)
    }
    if (multi && p < 8u) {              // ... and so have the exchange headers (the frame kernel packs into them again)
        SL.send[p >> 2].hdr[p & 3u] = 0u;
    }
    bool slow = false;
    uint32_t d = 0u;
    if (p < n) {
        const uint2 cr = A.cellrank[p];
        if (cr.x != INVALID_CELL) {   // ghosts / migrated agents / unused slots are dropped
            d = F.start[cr.x] + cr.y;
            const float4 pos = A.pos4n[p];
            const uint2 key = A.key2n[p];
            A.pos4[d] = pos;
            A.key2[d] = key;
            A.src[d] = p;
            const bool ghost = (key.y >> 2) & 1u;
            if (do_walk && !ghost && (P.walk_all || (key.y & 3u) == RTS_MOVING)) {
                slow = true;
                const uint32_t hint = A.trin[p];
                if (hint < M.n_tris && fabsf(pos.x) <= 1.0e9f && fabsf(pos.y) <= 1.0e9f) {
                    const iv2 q{(int32_t)pos.x, (int32_t)pos.y};   // static_cast<Vertex>(Vector2f): truncation
                    iv2 a, b, c;
                    int4 w1;
                    load_tri_verts(M, hint, a, b, c, w1);
                    // 64-bit products: identical to the reference's int32 wherever that does not overflow
                    const float d1 = tri_sign<true>(q, a, b), d2 = tri_sign<true>(q, b, c), d3 = tri_sign<true>(q, c, a);
                    // ... in a triangle of the map proper. The fan that ties the map's border to the super-triangle is not
                    // trusted: "strictly inside => in no other triangle" needs a planar triangulation, and the fans of a map
                    // stitched from tiles (scenario.tile_map) overlap each other and the map. Agents out there (pushed over
                    // the map edge) are always searched the reference's way.
                    const int W = (int)(P.map_nx * P.map_cell_size), Hh = (int)(P.map_ny * P.map_cell_size);
                    const bool proper = min(min(a.x, b.x), c.x) >= 0 && max(max(a.x, b.x), c.x) <= W &&
                                        min(min(a.y, b.y), c.y) >= 0 && max(max(a.y, b.y), c.y) <= Hh;
                    slow = !(proper && ((d1 > 0 && d2 > 0 && d3 > 0) || (d1 < 0 && d2 < 0 && d3 < 0)));
                }
            }
        }
    }
    // warp-aggregated append to the walk queue
    const unsigned need = __ballot_sync(0xffffffffu, slow);
    if (need) {
        const unsigned lane = threadIdx.x & 31u;
        uint32_t base = 0u;
        if (lane == (unsigned)(__ffs(need) - 1)) base = atomicAdd(A.walk_n + parity, (uint32_t)__popc(need));
        base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
        if (slow) A.walk_q[base + (uint32_t)__popc(need & ((1u << lane) - 1u))] = d;
    }
}

// The second half of the triangle walk (Triangulation::findTriangle): the search for the agents k_scatter queued. It
// writes the previous-order `tri` array the next frame carries along and only reads what the next frame kernel also
// only reads (pos4, src). A failed walk stores 0xFFFFFFFF.
// The reference's findTriangle for the agents k_scatter queued: CTA `cta` of `n_ctas` (threads_per_cta threads each)
// strides over the queue. Called by the leading CTAs of the NEXT frame kernel (the walk then overlaps the frame without a
// second stream: a few agents outside the map need 100+ dependent hops, ~40 us for one thread) and by k_walk.
template <bool WIDE>
__device__ __forceinline__ void drain_walk_queue(const RtsParams& P, const MapDev& M, const AgentArrays& A, int parity, uint32_t cta,
                                                 uint32_t n_ctas, uint32_t threads_per_cta, uint32_t t_in) {
    const uint32_t qn = A.walk_n[parity];
    for (uint32_t base = cta * threads_per_cta; base < qn; base += n_ctas * threads_per_cta) {
        const uint32_t i = base + t_in;
        const unsigned mask = __ballot_sync(0xffffffffu, i < qn);
        if (i < qn) {
            const uint32_t d = A.walk_q[i];
            const float4 pos = A.pos4[d];
            const uint32_t sp = A.src[d];
            uint32_t flags = 0;
            uint32_t t;
            // the int32 walk needs |q| <= 16383 as well as the table bound checked at upload
            const bool small = fabsf(pos.x) <= 16383.f && fabsf(pos.y) <= 16383.f;
            if (!WIDE && __all_sync(mask, small)) t = find_triangle_warp<false>(M, P, true, pos.x, pos.y, flags, mask);
            else t = find_triangle_warp<true>(M, P, true, pos.x, pos.y, flags, mask);
            *((volatile uint32_t*)A.tri + sp) = t;   // (the frame kernel may be reading it: see finish_agent)
        }
    }
}
template <bool WIDE>
__global__ void __launch_bounds__(256) k_walk(RtsParams P, MapDev M, AgentArrays A, int parity) {
    drain_walk_queue<WIDE>(P, M, A, parity, blockIdx.x, gridDim.x, blockDim.x, threadIdx.x);
}

// Slab exchange, receiving side: records that arrived from direction `dir` are appended behind the slots of the
// frame kernel (base = first slot of this direction's range), migrants as owned agents, halo records as ghosts.
// rts_slab_row_counts: owned agents per physics-grid row (the load figure slab rebalancing works with)
__global__ void k_row_hist(const uint32_t* __restrict__ n_dev, const uint2* __restrict__ key2, uint32_t* __restrict__ hist, uint32_t ns_ny) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= *n_dev) return;
    const uint32_t ky = key2[d].y;
    if ((ky >> 2) & 1u) return;   // ghost
    const uint32_t row = (uint32_t)cs_cy(ky);
    if (row < ns_ny) atomicAdd(hist + row, 1u);
}

// Peer exchange after k_halo_seed (no frame, hence no k_step_dense to do it): publish what was packed

// Slab exchange, receiving side: records that arrived from direction `dir` are appended behind the slots of the
// frame kernel (base = first slot of this direction's range), migrants as owned agents, halo records as ghosts.
// Peer mode: waits for the neighbour's flag of this epoch first (bounded: a neighbour that never delivers is reported
// through SL.overflow = 3 after ~20 s instead of hanging the device).
__global__ void k_ingest_remote(uint32_t base0, SlabDev SL, FineGrid F, AgentArrays A) {
    pdl_wait();
    pdl_trigger();
    if (SL.peer_mode && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) publish_to_peers(SL);   // (before this CTA waits below)
    const uint32_t cap_mig = SL.cap_mig, cap_halo = SL.cap_halo;
    const int dir = (int)blockIdx.y;                       // both directions in one launch
    ExBuf B = SL.recv[dir];
    __shared__ unsigned long long s_word;
    uint32_t peer_nm = 0u, peer_nh = 0u;
    if (SL.peer_mode) {
        const uint32_t e = *((volatile uint32_t*)SL.epoch);
        if (SL.has[dir]) {
            if (threadIdx.x == 0) {
                const volatile unsigned long long* flag = SL.flag_local[dir];
                unsigned long long t0 = 0, now = 0, w;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                // (24-bit epochs: "arrived" = not behind this frame's epoch + 1, modulo 2^24)
                // (acquire loads at system scope: everything the neighbour wrote before it released the word — fence + store in
                //  publish_to_peers — is visible behind the load that sees it, and through the barrier below to the whole CTA;
                //  no fence of its own: 2 us per frame less than volatile polling followed by a system-scope fence)
                auto poll = [&]() { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); return v; };
                while ((((uint32_t)((w = poll()) >> 40) - (e + 1u)) & 0xFFFFFFu) >= 0x800000u) {
                    __nanosleep(20);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (now - t0 > 20000000000ull) { *SL.overflow = 3u; w = 0ull; break; }
                }
                s_word = w;
            }
            __syncthreads();
            peer_nm = (uint32_t)(s_word >> 20) & 0xFFFFFu;
            peer_nh = (uint32_t)s_word & 0xFFFFFu;
        }
        B = carve_exbuf(SL.recv_base[dir][e & 1u], cap_mig, cap_halo);
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap_mig + cap_halo) return;
    const uint32_t slot = base0 + (uint32_t)dir * (cap_mig + cap_halo) + i;
    const bool is_mig = i < cap_mig;
    const uint32_t k = is_mig ? i : i - cap_mig;
    const uint32_t cnt = SL.peer_mode ? (is_mig ? min(peer_nm, cap_mig) : min(peer_nh, cap_halo))   // (0 without a neighbour)
                                      : (is_mig ? min(B.hdr[0], cap_mig) : min(B.hdr[1], cap_halo));
    if (k >= cnt) {
        A.cellrank[slot] = make_uint2(INVALID_CELL, 0u);
        return;
    }
    float4 pos;
    if (is_mig) {
        pos = B.m_pos4[k];
        A.pos4n[slot] = pos;
        A.key2n[slot] = B.m_key2[k];
        A.vel4n[slot] = B.m_vel4[k];
        A.nav4n[slot] = B.m_nav4[k];
        A.trin[slot] = B.m_tri[k];
        if (A.vel_outn) A.vel_outn[slot] = B.m_velout[k];
    } else {
        pos = B.h_pos4[k];
        A.pos4n[slot] = pos;
        A.key2n[slot] = B.h_key2[k];   // ghost bit set by the sender
        A.vel4n[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
        A.nav4n[slot] = B.h_nav4[k];   // (read by k_arrive only: a ghost is never stepped)
        A.trin[slot] = 0xFFFFFFFFu;
        if (A.vel_outn) A.vel_outn[slot] = make_float2(0.f, 0.f);
    }
    int fx, fy;
    fine_coords(F, pos.x, pos.y, fx, fy);
    const uint32_t cell = (uint32_t)fy * (uint32_t)F.nx + (uint32_t)fx;
    A.cellrank[slot] = make_uint2(cell, atomicAdd(F.count + cell, 1u));
}

// Slab exchange without a frame (after an upload): current state -> "n" buffers (ghosts dropped), halo bands packed.
__global__ void k_halo_seed(uint32_t n_bound, RtsParams P, FineGrid F, AgentArrays A, SlabDev SL) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *A.n_dev;
    const unsigned valid = __ballot_sync(0xffffffffu, p < n);
    if (p >= n) {
        if (p < n_bound) A.cellrank[p] = make_uint2(INVALID_CELL, 0u);
        return;
    }
    const float4 pos = A.pos4[p];
    const uint2 key = A.key2[p];
    const bool ghost = (key.y >> 2) & 1u;
    const unsigned active = __ballot_sync(valid, !ghost);
    if (ghost) {
        A.cellrank[p] = make_uint2(INVALID_CELL, 0u);
        return;
    }
    const uint32_t sp = A.src[p];
    A.pos4n[p] = pos;
    A.key2n[p] = key;
    A.vel4n[p] = A.vel4[sp];
    A.nav4n[p] = A.nav4[sp];
    A.trin[p] = A.tri[sp];
    if (A.vel_outn) A.vel_outn[p] = A.vel_out[sp];
    if (SL.has[0] && pos.y < SL.y_lo + SL.halo) {
        const uint32_t k = atomicAdd(SL.send[0].hdr + 1, 1u);
        const ExBuf B = pack_target(SL, 0);
        if (k < SL.cap_halo) { B.h_pos4[k] = pos; B.h_key2[k] = make_uint2(key.x, key.y | 4u); B.h_nav4[k] = A.nav4[sp]; }
        else *SL.overflow = 1u;
    }
    if (SL.has[1] && pos.y >= SL.y_hi - SL.halo) {
        const uint32_t k = atomicAdd(SL.send[1].hdr + 1, 1u);
        const ExBuf B = pack_target(SL, 1);
        if (k < SL.cap_halo) { B.h_pos4[k] = pos; B.h_key2[k] = make_uint2(key.x, key.y | 4u); B.h_nav4[k] = A.nav4[sp]; }
        else *SL.overflow = 1u;
    }
    if (SL.peer_mode && ((SL.has[0] && pos.y < SL.y_lo + SL.halo) || (SL.has[1] && pos.y >= SL.y_hi - SL.halo))) __threadfence_system();
    int fx, fy;
    fine_coords(F, pos.x, pos.y, fx, fy);
    const uint32_t cell = (uint32_t)fy * (uint32_t)F.nx + (uint32_t)fx;
    A.cellrank[p] = make_uint2(cell, warp_aggregated_rank(F.count, cell, active));
}

// Arrival rule (SeekSystem::finalizeSeek, SeekSystem.cpp:235-275) with frame-start snapshot semantics; runs before the
// frame kernel. A MOVING agent on its last leg (r_next == path_end) that is within 2*radius + sqrt(area/pi) of path_end
// stops, and stops the same-group members of its seek neighbour list (|dr|^2 < 10 (ri+rj)^2, Strategy.cpp:109-111) that
// are within finish_radius. Every stop adds pi*RHARD^2 to the group's finished area exactly once (atomicExch claim); the
// increments are all equal, so the float sum does not depend on their order.
__global__ void k_area_sync(MapDev M) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < M.n_groups) M.area_read[g] = M.area_acc[g];
}
// ... on a slab handle: the stops every rank has claimed up to the last exchange, added one by one
__global__ void k_area_sync_counts(MapDev M, float area_unit) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= M.n_groups) return;
    const uint32_t total = M.stop_global[g];
    uint32_t done = M.stop_applied[g];
    float area = M.area_read[g];
    for (; done < total; ++done) area += area_unit;
    M.area_read[g] = area;
    M.stop_applied[g] = total;
}

// Slab handles (multi): an agent near a slab boundary has neighbours on both sides. Every rank evaluates the rule for the
// agents it owns AND for its ghosts (halo records carry the navigation record, and the halo band is finish_radius wide
// while the rule is on: every agent that can stop one of this rank's agents is present here, with the same position,
// waypoint state and — the counts being all-reduced every frame — the same finished area as on its owner), and stops only
// agents it owns: each agent is stopped, and counted, exactly once, by its owner.
__global__ void k_arrive(uint32_t n_bound, RtsParams P, MapDev M, FineGrid F, AgentArrays A, int multi) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = multi ? *A.n_dev : n_bound;
    if (p >= n) return;
    const uint2 key = A.key2[p];
    const bool ghost_i = (key.y >> 2) & 1u;
    if ((key.y & 3u) != RTS_MOVING || (ghost_i && !multi)) return;
    const float4 nav4 = A.nav4[A.src[p]];
    const int32_t pid = __float_as_int(nav4.z);
    if (pid < 0 || (uint32_t)pid >= M.n_paths) return;   // (a stale id beyond the uploaded table counts as "no path")
    const uint32_t p0 = __ldg(M.path_off + pid), last = __ldg(M.path_off + pid + 1) - 1;
    const uint32_t w0 = min(p0 + (__float_as_uint(nav4.w) & 0xFFFFu), last);
    v2 r_next = V(nav4.x, nav4.y);
    if (r_next.x != r_next.x) { const float4 wa = __ldg(M.waypoints + 2 * (size_t)w0); r_next = V(wa.x, wa.y); }
    const float2 pe = __ldg(M.path_end + pid);
    const v2 path_end = V(pe.x, pe.y);
    if (!vequal(path_end, r_next)) return;
    const int32_t g_i = __ldg(M.path_group + pid);
    float4 me = A.pos4[p];
    me.z = fabsf(me.z);   // (sign bits of radius / mass carry the state, see pack_pos)
    const v2 r = V(me.x, me.y);
    const float finish_radius = sqrtf(M.area_read[g_i] / 3.14159265358979323846f);
    const float dist_to_end = sqrtf(dist2(path_end, r));
    if (!(dist_to_end < 2 * me.z + finish_radius)) return;
    const float area_unit = 3.14159265358979323846f * P.rhard * P.rhard;
    auto claim = [&](uint32_t q) {   // stop agent q (owned by this handle); the first claim of the frame counts it
        if (atomicExch(A.stop_req + q, 1u) == 0u) {
            if (multi) atomicAdd(M.stop_local + g_i, 1u);
            else atomicAdd(M.area_acc + g_i, area_unit);
        }
    };
    if (!ghost_i) claim(p);
    // neighbours: every fine cell that can hold an agent within finish_radius
    int fx, fy;
    fine_coords(F, r.x, r.y, fx, fy);
    const int reach = (int)ceilf(P.finish_radius * F.inv_cs);
    const int xlo = max(fx - reach, 0), xhi = min(fx + reach, F.nx - 1);
    for (int yy = max(fy - reach, 0); yy <= min(fy + reach, F.ny - 1); ++yy) {
        const uint32_t rowbase = (uint32_t)yy * (uint32_t)F.nx;
        const uint32_t qb = __ldg(F.start + rowbase + xlo), qe = __ldg(F.start + rowbase + xhi + 1);
        for (uint32_t q = qb; q < qe; ++q) {
            if (q == p) continue;
            const float4 o = A.pos4[q];
            const v2 dr = V(o.x, o.y) - r;
            const float rr = me.z + fabsf(o.z);
            if (!(norm2(dr) < 10.f * rr * rr)) continue;
            if (!(sqrtf(norm2(r - V(o.x, o.y))) < P.finish_radius)) continue;
            const uint2 ok = A.key2[q];
            if ((ok.y & 3u) != RTS_MOVING || ((ok.y >> 2) & 1u)) continue;   // (a ghost is stopped by its owner)
            const int32_t pj = __float_as_int(A.nav4[A.src[q]].z);
            if (pj < 0 || (uint32_t)pj >= M.n_paths || __ldg(M.path_group + pj) != g_i) continue;
            claim(q);
        }
    }
}

// Per-neighbour force terms of applyForces (PhysicsSystem.cpp:69-102): computed for one neighbour, then added to the
// running sums in the reference's visit order. Split in two so that a warp can compute the terms of many neighbours in
// parallel and still add them up in order (k_step_dense).
struct ForceSums {
    v2 repulsion, push, dr_nn;
    float n_neighbours;
    v2 align;          // alignment extension (RtsParams.enable_alignment): sum of v_j - v_i, and how many
    float n_align;
};
struct ForceTerms {
    v2 rep, push, dr;
    uint32_t mask;   // bit 0: repulsion term present, bit 1: push term, bit 2: scatter contribution, bit 3: alignment contribution
};
__device__ __forceinline__ ForceTerms neighbour_terms(v2 r, float radius, float mass_i, int state_i, float4 o, float rscatter2, float ralign2 = 0.f) {
    ForceTerms t;
    t.mask = 0u;
    t.rep = V(0.f, 0.f);
    t.push = V(0.f, 0.f);
    const v2 dr = V(o.x, o.y) - r;
    t.dr = dr;
    const int state_j = pos_state(o);
    const float r_collision = fabsf(o.z) + radius;
    const float mass_j = fabsf(o.w);
    const float r_collision_sq = r_collision * r_collision;
    const float d_surfs = norm2(dr);
    const float x = r_collision_sq / d_surfs;
    if (x > 0.8f) {
        const float m3 = 3.f * x * x * x;
        const float mag = (m3 < 50000.0f) ? m3 : 50000.0f;
        t.rep = (((-dr) / norm(dr)) * mag * mass_j) / (mass_i + mass_j);
        t.mask |= 1u;
    }
    if (state_j == RTS_MOVING && state_i == RTS_STANDING && x > 0.75f) {
        t.push = ((-dr) * x * mass_j) / (mass_i + mass_j);
        t.mask |= 2u;
    }
    if (state_j == RTS_MOVING && x > r_collision_sq / rscatter2) t.mask |= 4u;
    if (ralign2 > 0.f && state_i == RTS_MOVING && state_j == RTS_MOVING && x > r_collision_sq / ralign2) t.mask |= 8u;
    return t;
}
__device__ __forceinline__ void add_terms(ForceSums& S, v2 rep, v2 push, v2 dr, uint32_t mask) {
    if (mask & 1u) S.repulsion = S.repulsion + rep;
    if (mask & 2u) S.push = S.push + push;
    if (mask & 4u) {
        S.dr_nn = S.dr_nn + dr;
        S.n_neighbours += 1.f;
    }
}

// Everything of the frame after the neighbour sums: scatter term, clamp, decay, wall pass, seek, merge, wall pass,
// integration, migration / halo packing, and the write of the new record at slot p. Returns the fine cell of the new
// position (the caller ranks it).
// FAST (tolerance mode): the normalisations use rsqrt.approx instead of IEEE sqrt + division. Everything that DECIDES
// something — the wall contact set, the acos-table bin of turnTowards, the waypoint predicates of needsUpdating, the cell
// indices of the new position — is still evaluated with the reference's own arithmetic (or a shortcut that provably
// gives the same answer), so after one frame from the same state the discrete outputs (state, waypoint, r_next, flags,
// angle, angle_vel) are bit-identical to the exact mode and only vel / inertia / r carry rounding-level differences.
template <bool KEEP_VEL, bool MULTI, bool FAST, bool ALIGN = false>
__device__ __forceinline__ uint32_t finish_agent(uint32_t p, uint32_t sp, float4 me, uint2 mykey, float4 vel4, float4 nav4, const ForceSums& sums,
                                                 const RtsParams& P, const MapDev& M, const FineGrid& F, const AgentArrays& A, const SlabDev& SL) {
    const v2 r = V(me.x, me.y);
    const float radius = fabsf(me.z), mass_i = fabsf(me.w);
    const uint32_t gid = mykey.x;
    const int state_i = (int)(mykey.y & 3u);
    const uint32_t misc = __float_as_uint(nav4.w);
    const RtsUnitType ut = M.unit_types[(misc >> 16) & 0xFFu];
    v2 repulsion = sums.repulsion, push = sums.push, dr_nn = sums.dr_nn, scatter = V(0.f, 0.f);
    const float n_neighbours = sums.n_neighbours;

    // ---- P2 tail: applyForces, PhysicsSystem.cpp:104-116
    if (FAST) {
        // mean / |mean| = sum / |sum|
        if (n_neighbours > 1.f) scatter = dr_nn * (rsqrt_nr(norm2(dr_nn)) * ut.max_speed);
    } else {
        dr_nn = dr_nn / n_neighbours;
        if (n_neighbours > 1.f) scatter = (dr_nn / norm(dr_nn)) * ut.max_speed;
    }
    v2 inertia = V(vel4.x, vel4.y);
    inertia = inertia + ((repulsion * P.mul_repulse + push * P.mul_push) + scatter * P.mul_scatter);
    if (ALIGN && sums.n_align > 0.f) inertia = inertia + (sums.align / sums.n_align) * P.mul_align;   // extension, see RtsParams.enable_alignment

    // ---- P3: PhysicsSystem::update body, PhysicsSystem.cpp:22-44 (component velocity is 0 on entry, so the
    //      `vel /= speed*max_speed` clamp at :27-30 never fires)
    v2 vp = inertia;
    const float max_vel = ut.max_speed * P.mul_velocity;
    if (FAST) {
        const float s2 = norm2(vp);
        if (s2 > max_vel * max_vel) vp = vp * (max_vel * rsqrt_nr(s2));
    } else {
        vp = V(0.f, 0.f);
        float speed = norm(vp);
        if (speed > P.sys_max_speed) vp = vp / (speed * P.sys_max_speed);
        vp = vp + inertia;
        speed = norm(vp);
        if (speed > max_vel) vp = vp * (max_vel / speed);
    }
    const float lm = ut.lambda * P.mul_decay;
    const float decay = (lm < 1.f) ? lm : 1.f;
    inertia = inertia - inertia * decay;

    // ---- P4: wall pass #1, PhysicsSystem.cpp:130-162
    int n_walls = 0;
    if (may_touch_wall(M, r, radius))
        n_walls = contact_walls(M, P, r, radius, [&](float tx, float ty, uint32_t) { avoid_one(tx, ty, vp, inertia); });

    // ---- S1: SeekSystem::update per-agent body, SeekSystem.cpp:76-111
    v2 vs = V(0.f, 0.f);
    float angle = vel4.z, angle_vel = vel4.w;
    uint32_t wp = misc & 0xFFFFu;
    uint32_t flags = (misc >> 28) & 0xFu;
    v2 r_next = V(nav4.x, nav4.y);
    int32_t pid = __float_as_int(nav4.z);
    int state_out = state_i;
    if (P.enable_arrival && A.stop_req[p]) {   // stopSeeking, SeekSystem.cpp:136-144: no seek velocity, out of the group
        A.stop_req[p] = 0u;
        state_out = RTS_STANDING;
        angle_vel = 0.f;
        pid = -1;
    }
    if (state_out == RTS_MOVING && pid >= 0 && (uint32_t)pid < M.n_paths) {   // (a stale id beyond the uploaded table counts as "no path")
        const float4 hdr = __ldg(M.path_hdr + pid);
        const uint32_t p0 = __float_as_uint(hdr.x), p1 = __float_as_uint(hdr.y);
        const uint32_t last = p1 - 1;
        const uint32_t w0 = min(p0 + wp, last);
        const uint32_t w1 = min(w0 + 1, last);
        const float4 wa = __ldg(M.waypoints + 2 * (size_t)w0), wb = __ldg(M.waypoints + 2 * (size_t)w0 + 1);
        const float4 wn = __ldg(M.waypoints + 2 * (size_t)w1);
        if (r_next.x != r_next.x) r_next = V(wa.x, wa.y);
        const v2 r_nn = V(wn.x, wn.y);
        const v2 path_end = V(hdr.z, hdr.w);

        const v2 dr_to_target = r_next - r;
        const float n2_target = norm2(dr_to_target);
        // |dr_to_target| feeds three comparisons (against radius and RHARD) that decide discrete things. FAST: the
        // rsqrt-based norm (relative error < 3e-7) is replaced by the IEEE one whenever it is close to a threshold.
        float norm_dr, inv_norm_dr = 0.f;
        if (FAST) {
            inv_norm_dr = rsqrt_nr(n2_target);
            norm_dr = n2_target * inv_norm_dr;
            if (!(fabsf(norm_dr - radius) > 1e-5f * radius && fabsf(norm_dr - P.rhard) > 1e-5f * P.rhard)) norm_dr = sqrtf(n2_target);   // also 0 and NaN
        } else norm_dr = sqrtf(n2_target);
        {   // turnTowards, SeekSystem.cpp:344-385 (the component-local angle edit is never communicated, :322-342)
            float ang = angle;
            if (ang > 180) ang -= 360;
            if (ang < -180) ang += 360;
            float desired;
            if (n2_target == 0) desired = ut.desired_angle;
            else if (FAST) {
                // AcosTable<1000>::angle (core.h:139-150): bin = floor((x/|v| + 1) / 2 * 1000). The approximate bin
                // coordinate is within 1e-3 of the reference's; unless it is that close to an integer both floor alike
                const float bq = (dr_to_target.x * inv_norm_dr + 1.0f) * 500.f;
                const float bf = floorf(bq);
                const float bd = bq - bf;
                if (bd > 2e-3f && bd < 1.f - 2e-3f) {
                    int bin = (int)bf;
                    if (bin >= 1000) bin = 999;
                    if (bin < 0) bin = 0;
                    desired = __ldg(M.acos_table + bin) * (2.f * (dr_to_target.y > 0) - 1.f);
                } else desired = table_angle(M.acos_table, dr_to_target);
            } else desired = table_angle(M.acos_table, dr_to_target);
            float d_angle = desired - ang;
            if (d_angle > 180) d_angle = -360 + d_angle;
            if (d_angle < -180) d_angle = 360 + d_angle;
            angle_vel = (d_angle < ut.turn_rate) ? d_angle : ut.turn_rate;
            if (d_angle < 0) angle_vel = (d_angle > -ut.turn_rate) ? d_angle : -ut.turn_rate;
            if (fabsf(d_angle) < ut.turn_rate) angle_vel = 0;
        }
        if (norm_dr > radius) vs = FAST ? dr_to_target * (inv_norm_dr * ut.seek_max_speed) : (dr_to_target / norm_dr) * ut.seek_max_speed;

        bool upd;  // needsUpdating, SeekSystem.cpp:181-233
        if (vequal(r_next, path_end)) upd = false;
        else {
            const v2 t = r_nn - r_next;
            const bool can_move = dot(-dr_to_target, t) > 0;
            upd = (norm_dr < P.rhard && can_move) || veq(r_next, r_nn);
            const v2 cp = closest_point_on_edge(r, V(wa.z, wa.w), V(wb.x, wb.y), wb.z);
            if (norm_dr < radius || dot(dr_to_target, t) < 0) upd = true;
            else {
                if (dist2(cp, r) < dist2(r, r_next) && wb.z > 0 && dist2(cp, r) < 16 * P.rhard * P.rhard && norm2(t) != 0) {
                    r_next = cp;
                    upd = false;
                } else if (vequal(r_next, r_nn) && !veq(r_next, path_end)) upd = true;
            }
        }
        if (upd) {  // updateTarget, SeekSystem.cpp:277-304
            const v2 dn = r_nn - r;
            r_next = r_nn;
            if (p0 + wp < last) wp++;
            if (FAST) {
                const float n2 = norm2(dn);
                if (n2 > 0) vs = dn * (rsqrt_nr(n2) * ut.seek_max_speed);
            } else {
                const float ndn = norm(dn);
                if (ndn > 0) vs = (dn / ndn) * ut.seek_max_speed;
            }
            if (!veq(r_nn, path_end)) flags |= RTS_FLAG_REPLAN;
        }
    }

    // ---- M: merge, wall pass #2 (same contact set: r has not moved), integrate (ECS.cpp:92-113)
    v2 v = (V(0.f, 0.f) + vp) + vs;
    if (n_walls > 0) contact_walls(M, P, r, radius, [&](float tx, float ty, uint32_t) { avoid_one(tx, ty, v, inertia); });
    const v2 r_new = r + v * P.dt;
    angle = angle + angle_vel;

    // ---- K7 (triangle walk on the new position) runs in k_scatter (hint test) + drain_walk_queue (search); agents that
    // are not walked keep their index. The index is carried forward for every agent: k_scatter uses it as its hint. The
    // search for the previous frame's queued agents may be running in this kernel's leading CTAs right now, so this read
    // sees either its result or the entry it is about to replace — the same agent's triangle one frame earlier, an
    // equally valid hint (hints are verified).
    A.trin[p] = *((const volatile uint32_t*)A.tri + sp);

    // ---- write the new record at this slot; bin it for the next order
    const float4 pos_new = pack_pos(r_new.x, r_new.y, radius, mass_i, (uint32_t)state_out);
    const uint32_t cx_new = axis_cell(r_new.x, P.ns_cell_size, M.inv_ns_cs, P.ns_nx), cy_new = axis_cell(r_new.y, P.ns_cell_size, M.inv_ns_cs, P.ns_ny);
    uint32_t cs_new = pack_cs(cx_new, cy_new, 0u, (uint32_t)state_out);
    const float4 vel_new = make_float4(inertia.x, inertia.y, angle, angle_vel);
    const float4 nav_new = make_float4(r_next.x, r_next.y, __int_as_float(pid), __uint_as_float(pack_misc(wp, (misc >> 16) & 0xFFu, (misc >> 24) & 0xFu, flags)));
    if (MULTI) {
        // ownership follows the physics-grid row of the new position (integer, identical on every rank)
        // (an agent whose position has just become NaN — coincident agents, 0/0 in the reference too — stays where it is)
        const int dir = !(fabsf(r_new.y) <= 3.0e38f) ? -1 : ((int)cy_new < SL.row_lo) ? 0 : (((int)cy_new >= SL.row_hi) ? 1 : -1);
        if (SL.interior_launch && ((dir >= 0 && SL.has[dir]) || (SL.has[0] && r_new.y < SL.y_lo + SL.halo) ||
                                   (SL.has[1] && r_new.y >= SL.y_hi - SL.halo)))
            *SL.overflow = 2u;   // the exchange of this frame is already under way: report (rts_sync), never drop silently
        if (dir >= 0 && SL.has[dir]) {
            // migrate: full record to the neighbour; the local copy stays one more frame as a ghost (it is within
            // one frame's travel of the boundary, i.e. inside the halo band of this slab)
            const ExBuf B = pack_target(SL, dir);
            const uint32_t k = atomicAdd(SL.send[dir].hdr + 0, 1u);
            if (k < SL.cap_mig) {
                B.m_pos4[k] = pos_new;
                B.m_key2[k] = make_uint2(gid, cs_new);
                B.m_vel4[k] = vel_new;
                B.m_nav4[k] = nav_new;
                B.m_tri[k] = 0xFFFFFFFFu;   // walked by the receiver's reorder pass
                B.m_velout[k] = make_float2(v.x, v.y);
            } else *SL.overflow = 1u;
            cs_new |= 4u;
        } else {
            if (SL.has[0] && r_new.y < SL.y_lo + SL.halo) {
                const uint32_t k = atomicAdd(SL.send[0].hdr + 1, 1u);
                const ExBuf B = pack_target(SL, 0);
                if (k < SL.cap_halo) { B.h_pos4[k] = pos_new; B.h_key2[k] = make_uint2(gid, cs_new | 4u); B.h_nav4[k] = nav_new; }
                else *SL.overflow = 1u;
            }
            if (SL.has[1] && r_new.y >= SL.y_hi - SL.halo) {
                const uint32_t k = atomicAdd(SL.send[1].hdr + 1, 1u);
                const ExBuf B = pack_target(SL, 1);
                if (k < SL.cap_halo) { B.h_pos4[k] = pos_new; B.h_key2[k] = make_uint2(gid, cs_new | 4u); B.h_nav4[k] = nav_new; }
                else *SL.overflow = 1u;
            }
        }
        // peer exchange: this thread's records must have ARRIVED in the neighbour's memory before the kernel may end (the
        // flag that announces them is raised by a later kernel, which orders nothing about another kernel's posted writes)
        if (SL.peer_mode && (dir >= 0 || (SL.has[0] && r_new.y < SL.y_lo + SL.halo) || (SL.has[1] && r_new.y >= SL.y_hi - SL.halo)))
            __threadfence_system();
    }
    A.pos4n[p] = pos_new;
    A.key2n[p] = make_uint2(gid, cs_new);
    A.vel4n[p] = vel_new;
    A.nav4n[p] = nav_new;
    if (KEEP_VEL) A.vel_outn[p] = make_float2(v.x, v.y);
    int nfx, nfy;
    fine_coords(F, r_new.x, r_new.y, nfx, nfy);
    return (uint32_t)nfy * (uint32_t)F.nx + (uint32_t)nfx;
}

// ------------------------------------------------------------------------------------------------
// Tolerance mode: the candidates of the three fine rows are visited as ONE virtual range (no per-row loops: the trip
// count of a warp is the longest total, not the sum of the longest rows) and the force terms of applyForces
// (PhysicsSystem.cpp:69-102) are evaluated branch-free for every candidate and added where they apply — no neighbour
// list, no ordering, one rsqrt.approx per candidate. The ACCEPTANCE test is the reference's own arithmetic
// (dx*dx + dy*dy < r_max2, separately rounded products), so the neighbour set is the reference's set.
struct FastScan {
    v2 r;
    float radius, mass_i;
    bool standing;        // state_i == STANDING: the push term applies
    float r_max2, rscatter2;
    uint32_t self;        // own slot
    bool near_edge;       // the 3x3 window touches a clamp cell of the fine grid: candidates may lie outside the map
    int cxi, cyi;         // own physics-grid cell
    uint32_t qb0, len0, qb1, len1, qb2, len2;   // slot ranges of the three fine rows
    // alignment extension
    bool moving;          // state_i == MOVING
    float ralign2;
    v2 vi;                // own merged velocity of the previous frame
    const uint32_t* src;  // slot -> previous-order index
    const float2* vel_out;
};
// PUSH: some lane of the warp is STANDING (the push term exists). CHECKS: some lane's window touches the edge of the
// fine grid (physics-cell adjacency must be checked, see cell_is_adjacent) or RSCATTER^2 < r_max2 (the scatter condition
// is not implied by the acceptance). Both are decided per warp, so that the common case carries neither.
template <bool PUSH, bool CHECKS, bool ALIGN>
__device__ __forceinline__ void fast_candidate(ForceSums& S, const FastScan& c, const float4 o, const uint32_t q, bool valid, const uint2* __restrict__ key2) {
    const float dx = o.x - c.r.x, dy = o.y - c.r.y;
    const float d2 = dx * dx + dy * dy;                 // norm2(dr), as the reference rounds it
    bool acc = valid && d2 < c.r_max2 && q != c.self;
    if (CHECKS) {
        if (c.near_edge && acc) acc = cell_is_adjacent(key2[q].y, c.cxi, c.cyi);
    }
    const float mass_j = fabsf(o.w);
    const float rc = fabsf(o.z) + c.radius;
    const float inv_d = rsqrt_fast(d2);
    const float u = rc * inv_d;
    const float x = u * u;                               // r_collision^2 / |dr|^2
    const float w = mass_j * rcp_fast(c.mass_i + mass_j);
    // min(50000, 3 x^3) = 3 min(50000/3, x^3): the factor 3 is applied once, to the sum (fast_accumulate)
    const float mag = fminf((x * x) * x, 16666.666f);
    const bool rep_on = acc & (x > 0.8f);
    const float s_rep = rep_on ? (mag * w) * inv_d : 0.f;
    S.repulsion.x = __fmaf_rn(-dx, s_rep, S.repulsion.x);
    S.repulsion.y = __fmaf_rn(-dy, s_rep, S.repulsion.y);
    // state_j == MOVING (= 0): both sign bits of {radius, mass} clear
    const bool moving_j = acc & ((__float_as_int(o.z) | __float_as_int(o.w)) >= 0);
    if (PUSH) {
        const bool push_on = c.standing & moving_j & (x > 0.75f);
        const float s_push = push_on ? x * w : 0.f;
        S.push.x = __fmaf_rn(-dx, s_push, S.push.x);
        S.push.y = __fmaf_rn(-dy, s_push, S.push.y);
    }
    // x > rc2 / RSCATTER^2  <=>  d2 < RSCATTER^2, implied by d2 < r_max2 unless RSCATTER^2 < r_max2
    const bool sc = CHECKS ? (moving_j & (d2 < c.rscatter2)) : moving_j;
    const float m = sc ? 1.f : 0.f;
    S.dr_nn.x = __fmaf_rn(dx, m, S.dr_nn.x);             // (exact: m is 0 or 1)
    S.dr_nn.y = __fmaf_rn(dy, m, S.dr_nn.y);
    S.n_neighbours += m;
    if (ALIGN) {
        if (c.moving && moving_j && d2 < c.ralign2) {
            const float2 vj = c.vel_out[c.src[q]];
            S.align.x += vj.x - c.vi.x;
            S.align.y += vj.y - c.vi.y;
            S.n_align += 1.f;
        }
    }
}
// candidates v0, v0 + vstep, ... of the virtual range (vstep = 1: one thread does them all; vstep = lanes of a group)
template <bool PUSH, bool CHECKS, bool ALIGN>
__device__ __forceinline__ void fast_loop(ForceSums& S, const FastScan& c, const float4* __restrict__ pos4, const uint2* __restrict__ key2,
                                          uint32_t v0, uint32_t vstep) {
    const uint32_t l01 = c.len0 + c.len1, total = l01 + c.len2;
    const uint32_t o1 = c.qb1 - c.len0, o2 = c.qb2 - l01;
    auto slot = [&](uint32_t v) { return v + (v < c.len0 ? c.qb0 : (v < l01 ? o1 : o2)); };
#pragma unroll 1
    for (uint32_t v = v0; v < total; v += 2u * vstep) {
        const uint32_t vb = v + vstep;
        const bool hb = vb < total;
        const uint32_t qa = slot(v), qb = slot(hb ? vb : v);
        const float4 oa = pos4[qa], ob = pos4[qb];       // both loads in flight
        fast_candidate<PUSH, CHECKS, ALIGN>(S, c, oa, qa, true, key2);
        fast_candidate<PUSH, CHECKS, ALIGN>(S, c, ob, qb, hb, key2);
    }
}
// `lanes`: the lanes of the warp that call this together (the variant is chosen per warp)
template <bool ALIGN>
__device__ __forceinline__ void fast_accumulate(ForceSums& S, const FastScan& c, const float4* __restrict__ pos4, const uint2* __restrict__ key2,
                                                uint32_t v0, uint32_t vstep, unsigned lanes) {
    if (ALIGN) fast_loop<true, true, true>(S, c, pos4, key2, v0, vstep);   // (the opt-in extension takes the general loop)
    else {
        const bool push = __any_sync(lanes, c.standing);
        const bool checks = __any_sync(lanes, c.near_edge) || !(c.rscatter2 >= c.r_max2);
        if (!push && !checks) fast_loop<false, false, false>(S, c, pos4, key2, v0, vstep);
        else if (push && !checks) fast_loop<true, false, false>(S, c, pos4, key2, v0, vstep);
        else if (!push) fast_loop<false, true, false>(S, c, pos4, key2, v0, vstep);
        else fast_loop<true, true, false>(S, c, pos4, key2, v0, vstep);
    }
    S.repulsion = S.repulsion * 3.f;   // see fast_candidate: the sum was built with min(50000/3, x^3)
}

// ------------------------------------------------------------------------------------------------
// The fused frame. One thread per slot (agents of one fine cell are adjacent, so a warp reads a few
// neighbouring cells and the neighbour records come out of L1).
// FAST = false: bit-exact mode — accepted neighbours are collected, sorted into the reference's visit order and added
// up in that order with IEEE arithmetic. FAST = true: tolerance mode (see fast_accumulate).
// PHASE 0: every slot; PHASE 1: the edge rows of a slab only (see SlabDev), PHASE 2: everything else.
template <bool KEEP_VEL, bool MULTI, bool FAST, int PHASE = 0, bool ALIGN = false>
__global__ void __launch_bounds__(FAST ? FAST_BLOCK : STEP_BLOCK, FAST ? FAST_MIN_BLOCKS : STEP_MIN_BLOCKS) k_step(uint32_t n_bound, RtsParams P, MapDev M, FineGrid F, AgentArrays A, SlabDev SL,
                                                                                 uint32_t walk_ctas, int walk_parity, uint32_t n_tiles) {
    // Work is handed out in TILES. Default: a tile is a CTA's worth of slots, tile = blockIdx.x. PERSIST (build switch
    // RTS_PERSISTENT_STEP, launches that cover all slots): a tile is 32 slots, the grid is one residency of the device, and
    // every WARP takes tiles from a device-wide counter (A.work, reset by k_scatter) until none is left.
    // The first walk_ctas tiles finish the triangle walk of the previous frame (the agents its k_scatter queued).
    constexpr bool PERSIST = USE_PERSISTENT_STEP && PHASE == 0;
    constexpr uint32_t BLOCK = FAST ? FAST_BLOCK : STEP_BLOCK;
    constexpr uint32_t TILE = PERSIST ? 32u : BLOCK;   // slots per tile
    __shared__ uint2 s_kq[FAST ? 1 : NBR_CAP][FAST ? 1 : STEP_BLOCK];   // {key, candidate index} of the NBR_CAP smallest keys seen, ascending

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x;
    const uint32_t t_in = PERSIST ? ((uint32_t)tid & 31u) : (uint32_t)tid;   // index inside the tile
    const uint32_t n = MULTI ? *A.n_dev : n_bound;   // slabs: the population changes every frame, the count lives on the device
    // edge slots of a slab: [0, edge_a) and [edge_b, n)
    // (the first and the last fine row are always edge rows: positions outside the slab's grid are clamped into them,
    //  among them agents that left the map, whose wrapped physics row (Grid.cpp:18-24) can make them migrate)
    uint32_t edge_a = 0u, edge_b = 0xFFFFFFFFu;
    if (PHASE != 0) {
        edge_a = __ldg(F.start + (uint32_t)(SL.has[0] ? SL.edge_row_lo : 1) * (uint32_t)F.nx);
        edge_b = __ldg(F.start + (uint32_t)(SL.has[1] ? SL.edge_row_hi : F.ny - 1) * (uint32_t)F.nx);
        edge_b = max(edge_b, edge_a);
    }
    const uint32_t cover = SL.edge_blocks * BLOCK;   // slots the PHASE 1 grid reaches on either side
    const uint32_t warps_in_grid = gridDim.x * (BLOCK / 32u);
    uint32_t tile = PERSIST ? blockIdx.x * (BLOCK / 32u) + ((uint32_t)tid >> 5) : blockIdx.x;   // first tile: the warp's own number
    for (;;) {
    uint32_t next_raw = 0u;
    if (PERSIST) {
        if (tile >= n_tiles) break;
        // (ptxas turns an atomic add on a warp-uniform address into elect + atomic + SHUFFLE of the result — even for one
        //  active lane, even from inline PTX, even spelled atom.inc — and the shuffle waits for the round trip right here,
        //  while the result is only needed at the end of the loop body. The address below is A.work for every lane, but not
        //  provably so: tiles stay far below 2^25.)
        const uint32_t opaque_zero = (tile * 32u + t_in) >> 30;
        if (t_in == 0u) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(next_raw) : "l"(A.work + opaque_zero) : "memory");
    }
    if (tile < walk_ctas) {
        if (M.coords_fit_i32) drain_walk_queue<false>(P, M, A, walk_parity, tile, walk_ctas, TILE, t_in);
        else drain_walk_queue<true>(P, M, A, walk_parity, tile, walk_ctas, TILE, t_in);
    } else {
    const uint32_t bid = tile - walk_ctas;
    uint32_t block0 = bid * TILE;
    if (PHASE == 1 && bid >= SL.edge_blocks) block0 = edge_b + (bid - SL.edge_blocks) * BLOCK;
    const uint32_t p = block0 + t_in;
    // (slabs: the live count n is a word in device memory; the slot's record is requested for every slot of the handle's
    //  capacity — allocated, possibly stale beyond n — so that the two round trips overlap instead of following each other)
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 mykey = make_uint2(0u, 4u);
    if (p < n_bound) {
        me = A.pos4[p];
        mykey = A.key2[p];
    }
    const bool in_range = p < n;
    if (MULTI && !in_range) {
        me = make_float4(0.f, 0.f, 0.f, 0.f);
        mykey = make_uint2(0u, 4u);
    }
    const v2 r = V(me.x, me.y);
    const bool ghost = (mykey.y >> 2) & 1u;
    bool mine = true;   // does this launch step this slot?
    if (PHASE != 0) {
        const bool covered = (p < edge_a && p < cover) || (p >= edge_b && p - edge_b < cover);
        mine = (PHASE == 1) ? (covered && (bid < SL.edge_blocks ? p < edge_a : true)) : !covered;
    }
    const unsigned active = __ballot_sync(0xffffffffu, in_range && !ghost && mine);  // lanes that reach the re-bin at the end
    int fx, fy;
    fine_coords(F, r.x, r.y, fx, fy);

    if (!in_range) {
        if (MULTI && PHASE != 1 && p < n_bound) A.cellrank[p] = make_uint2(INVALID_CELL, 0u);   // unused slot
    } else if (!mine) {
    } else if (ghost) {  // halo copy: neighbour only, dropped from the next order
        A.cellrank[p] = make_uint2(INVALID_CELL, 0u);
    } else {
    const uint32_t sp = A.src[p];
    const float radius = fabsf(me.z), mass_i = fabsf(me.w);
    const int state_i = (int)(mykey.y & 3u);
    const int cxi = cs_cx(mykey.y), cyi = cs_cy(mykey.y);

    // ---- P1: neighbour pass (Strategy.cpp:10-68) over the 3x3 fine cells
    const int xlo = max(fx - 1, 0), xhi = min(fx + 1, F.nx - 1);
    const int ylo = max(fy - 1, 0), yhi = min(fy + 1, F.ny - 1);
    // a candidate outside the map sits in a clamp cell of the fine grid; only then can its physics-grid cell fail to be
    // one of the nine calcNearestCells visits (see cell_is_adjacent)
    const bool near_edge = fx <= 1 || fx >= F.nx - 2 || fy <= 1 || fy >= F.ny - 2;

    // candidate index ranges of the three fine rows (each row's 3 cells are contiguous in slot order)
    uint32_t row_qb[3], row_qe[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int yy = fy - 1 + k;
        if (yy >= ylo && yy <= yhi) {
            const uint32_t rowbase = (uint32_t)yy * (uint32_t)F.nx;
            row_qb[k] = __ldg(F.start + rowbase + xlo);
            row_qe[k] = __ldg(F.start + rowbase + xhi + 1);
        } else {
            row_qb[k] = 0u;
            row_qe[k] = 0u;
        }
    }
    ForceSums acc{V(0.f, 0.f), V(0.f, 0.f), V(0.f, 0.f), 0.f, V(0.f, 0.f), 0.f};
    const float rscatter2 = P.rscatter * P.rscatter;
    const float ralign2 = ALIGN ? P.ralign * P.ralign : 0.f;
    bool to_dense = false;
    uint32_t cell = INVALID_CELL;
    if (FAST) {
        FastScan c;
        if (ALIGN) {
            const float2 vi = A.vel_out[sp];
            c.moving = state_i == RTS_MOVING; c.ralign2 = ralign2; c.vi = V(vi.x, vi.y); c.src = A.src; c.vel_out = A.vel_out;
        }
        c.r = r; c.radius = radius; c.mass_i = mass_i; c.standing = state_i == RTS_STANDING;
        c.r_max2 = P.r_max2; c.rscatter2 = rscatter2; c.self = p; c.near_edge = near_edge; c.cxi = cxi; c.cyi = cyi;
        c.qb0 = row_qb[0]; c.len0 = row_qe[0] - row_qb[0];
        c.qb1 = row_qb[1]; c.len1 = row_qe[1] - row_qb[1];
        c.qb2 = row_qb[2]; c.len2 = row_qe[2] - row_qb[2];
        // An agent in a crowd would keep this thread busy long after the rest of its warp is done; it is queued for
        // k_step_dense, where a group of lanes shares its candidates.
        to_dense = USE_DENSE_KERNEL && !ALIGN && c.len0 + c.len1 + c.len2 > FAST_DENSE_CAND;   // (the alignment extension stays in this kernel)
        const unsigned lanes = __ballot_sync(active, !to_dense);
        if (to_dense) A.dense_q[atomicAdd(A.dense_n, 1u)] = p;
        else {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.vel4 + sp));   // the frame tail's loads: on their way while the candidates are scanned
            asm volatile("prefetch.global.L2 [%0];" ::"l"(A.nav4 + sp));
            fast_accumulate<ALIGN>(acc, c, A.pos4, A.key2, 0u, 1u, lanes);
            // (the rest of the agent's record is fetched only now: nothing of it is needed while the candidates are scanned)
            cell = finish_agent<KEEP_VEL, MULTI, true, ALIGN>(p, sp, me, mykey, A.vel4[sp], A.nav4[sp], acc, P, M, F, A, SL);
        }
    } else {
    const float4 vel4 = A.vel4[sp];
    const float4 nav4 = A.nav4[sp];
    //      accepted neighbours are kept sorted by (reference visit rank of their 60-unit cell, gid) = the reference's visit order
    const float4* cand_pos = A.pos4;
    const uint2* cand_key = A.key2;
    const uint32_t self_idx = p;
    auto for_each_accepted = [&](auto&& fn) {   // candidates within the interaction range, in slot order
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const uint32_t qe = row_qe[k];
#pragma unroll 1
            for (uint32_t q0 = row_qb[k]; q0 < qe; q0 += 4) {
                float4 o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) o[u] = cand_pos[min(q0 + u, qe - 1u)];   // 4 independent loads in flight
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t q = q0 + u;
                    const v2 dr = V(o[u].x, o[u].y) - r;
                    if (q < qe && norm2(dr) < P.r_max2 && q != self_idx && (!near_edge || cell_is_adjacent(cand_key[q].y, cxi, cyi))) fn(q);
                }
            }
        }
    };
    // ordering key of a neighbour: (visit rank of its 60-unit cell, gid)
    auto key_of = [&](uint32_t q) -> uint32_t {
        const uint2 ok = cand_key[q];
        const int dxc = cs_cx(ok.y) - cxi;
        const int dyc = cs_cy(ok.y) - cyi;
        // calcNearestCells order (Grid.cpp:103-118): dx outer, dy inner, centre appended last (Strategy.cpp:34)
        const int idx = (dxc + 1) * 3 + (dyc + 1);
        const uint32_t rank = (idx < 4) ? (uint32_t)idx : (idx == 4 ? 8u : (uint32_t)(idx - 1));
        return (rank << 28) | ok.x;
    };

    // applyForces loop body, PhysicsSystem.cpp:69-102
    v2 v_own = V(0.f, 0.f);
    if (ALIGN) { const float2 vi = A.vel_out[sp]; v_own = V(vi.x, vi.y); }
    auto accumulate = [&](uint32_t q) {
        const ForceTerms t = neighbour_terms(r, radius, mass_i, state_i, cand_pos[q], rscatter2, ralign2);
        add_terms(acc, t.rep, t.push, t.dr, t.mask);
        if (ALIGN && (t.mask & 8u)) {
            const float2 vj = A.vel_out[A.src[q]];
            acc.align = acc.align + (V(vj.x, vj.y) - v_own);
            acc.n_align += 1.f;
        }
    };

    // Accepted candidates beyond NBR_CAP are parked (index only) in a per-thread list in local memory.
    uint2 spill[USE_SPILL ? SPILL_CAP : 1];
    int cnt = 0;
    for_each_accepted([&](uint32_t q) {
        if (cnt < NBR_CAP) s_kq[cnt][tid].y = q;
        else if (USE_SPILL && cnt - NBR_CAP < SPILL_CAP) spill[cnt - NBR_CAP].y = q;
        ++cnt;
    });
    // A crowded agent (DENSE_MIN < neighbours <= DENSE_MAX) would keep this thread busy for tens of thousands of
    // instructions while the rest of its warp idles, and the slowest such thread sets the duration of the whole kernel.
    // It is queued for k_step_dense instead, where a warp shares the scan, the sort and the force terms of ONE agent.
    to_dense = USE_DENSE_KERNEL && !ALIGN && cnt > DENSE_MIN && cnt <= DENSE_MAX;   // (the alignment extension stays in this kernel: re-scan path)
    if (to_dense) {
        A.dense_q[atomicAdd(A.dense_n, 1u)] = p;
    } else {
    {   // keys + insertion sort of the shared-memory part, every lane in step
        const int m = min(cnt, NBR_CAP);
        for (int k = 0; k < m; ++k) s_kq[k][tid].x = key_of(s_kq[k][tid].y);
        for (int k = 1; k < m; ++k) {
            const uint2 e = s_kq[k][tid];
            int pos = k;
            while (pos > 0 && s_kq[pos - 1][tid].x > e.x) {
                s_kq[pos][tid] = s_kq[pos - 1][tid];
                --pos;
            }
            s_kq[pos][tid] = e;
        }
    }
    if (cnt <= NBR_CAP) {
        for (int k = 0; k < cnt; ++k) accumulate(s_kq[k][tid].y);
    } else {
        // Dense crowd. Shared memory keeps the NBR_CAP smallest keys, sorted; an entry that does not fit (pushed out, or
        // larger than everything kept) goes back to the spill list. The kept keys are accumulated in order; then the
        // spill — all of it larger than what was just used — is filtered through the same bounded insertion, NBR_CAP at
        // a time. Only crowds beyond NBR_CAP+SPILL_CAP neighbours re-scan the candidates (above the last used key).
        int kept = NBR_CAP, n_spill = 0;
        auto offer = [&](uint2 e, bool to_spill) {   // bounded sorted insertion; the loser goes to the spill list
            int pos;
            if (kept < NBR_CAP) pos = kept++;
            else {
                uint2 out = e;
                if (e.x < s_kq[NBR_CAP - 1][tid].x) { out = s_kq[NBR_CAP - 1][tid]; pos = NBR_CAP - 1; }
                else pos = -1;
                if (to_spill) spill[n_spill++] = out;   // n_spill never overtakes the read index: in-place compaction
                if (pos < 0) return;
            }
            while (pos > 0 && s_kq[pos - 1][tid].x > e.x) {
                s_kq[pos][tid] = s_kq[pos - 1][tid];
                --pos;
            }
            s_kq[pos][tid] = e;
        };
        if (USE_SPILL && cnt <= NBR_CAP + SPILL_CAP) {
            const int m0 = cnt - NBR_CAP;
            for (int k = 0; k < m0; ++k) {
                const uint32_t q = spill[k].y;
                offer(make_uint2(key_of(q), q), true);
            }
            for (int k = 0; k < kept; ++k) accumulate(s_kq[k][tid].y);
            while (n_spill > 0) {   // rounds over the spill only
                const int m = n_spill;
                n_spill = 0;
                kept = 0;
                for (int k = 0; k < m; ++k) offer(spill[k], true);
                for (int k = 0; k < kept; ++k) accumulate(s_kq[k][tid].y);
            }
        } else {
            // very dense crowd: the sorted list holds the NBR_CAP smallest of the FIRST NBR_CAP accepted only; rebuild it
            // over all candidates, then keep re-scanning above the last key used
            kept = 0;
            for_each_accepted([&](uint32_t q) { offer(make_uint2(key_of(q), q), false); });
            for (int k = 0; k < kept; ++k) accumulate(s_kq[k][tid].y);
            int done = kept;
            while (done < cnt) {
                const uint32_t floor_key = s_kq[kept - 1][tid].x;
                kept = 0;
                for_each_accepted([&](uint32_t q) {
                    const uint32_t key = key_of(q);
                    if (key > floor_key) offer(make_uint2(key, q), false);
                });
                for (int k = 0; k < kept; ++k) accumulate(s_kq[k][tid].y);
                done += kept;
            }
        }
    }

    cell = finish_agent<KEEP_VEL, MULTI, false, ALIGN>(p, sp, me, mykey, vel4, nav4, acc, P, M, F, A, SL);
    }
    }
    const uint32_t rank = warp_aggregated_rank(F.count, cell, active);
    if (!to_dense) A.cellrank[p] = make_uint2(cell, rank);
    }
    }   // (tile of slots)
    if (!PERSIST) break;
    {   // (volatile + memory clobber: the compiler must not move the shuffle — and with it the wait for the counter's round
        //  trip — up to the atomicAdd at the head of the loop body)
        uint32_t nxt;
        asm volatile("shfl.sync.idx.b32 %0, %1, 0, 0x1f, 0xffffffff;" : "=r"(nxt) : "r"(next_raw) : "memory");
        tile = warps_in_grid + nxt;
    }
    }   // (tiles)
}

// One lane GROUP per crowded agent (see k_step), DENSE_LANES = 16 lanes: two agents per warp. The kernel's duration is the
// number of crowded agents divided by the agents in flight (registers allow ~28 warps per SM) times one agent's chain of
// dependent memory round trips, so halving the group doubles the throughput; most crowded agents have 19..40 neighbours
// and fill a 16-lane group well. The lanes of a group scan the candidates together, sort the accepted neighbours by
// (visit rank, gid) with a bitonic sort in shared memory, compute the force terms of DENSE_LANES neighbours at a time, and
// add them up in order (lanes 0..6 of the group carry the seven sums); lane 0 of the group then finishes the frame for
// the agent. Loop bounds are made warp-uniform (maximum over the two groups) so that every __syncwarp is reached by all.
#ifndef RTS_DENSE_LANES
#define RTS_DENSE_LANES 32
#endif
constexpr uint32_t DENSE_LANES = RTS_DENSE_LANES;
constexpr uint32_t DENSE_GROUPS = 32u / DENSE_LANES;   // agents per warp
#ifndef RTS_DENSE_MIN_BLOCKS
#define RTS_DENSE_MIN_BLOCKS 8
#endif
template <bool KEEP_VEL, bool MULTI>
__global__ void __launch_bounds__(128, RTS_DENSE_MIN_BLOCKS) k_step_dense(RtsParams P, MapDev M, FineGrid F, AgentArrays A, SlabDev SL) {
    __shared__ uint2 s_list[4 * DENSE_GROUPS][DENSE_MAX];
    __shared__ uint2 s_sorted[4 * DENSE_GROUPS][DENSE_MAX];
    __shared__ float s_terms[4 * DENSE_GROUPS][7][DENSE_LANES];
    pdl_wait();
    pdl_trigger();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = lane / DENSE_LANES;        // which agent of the warp
    const unsigned gl = lane % DENSE_LANES;         // lane inside the group
    const unsigned gshift = grp * DENSE_LANES;
    const unsigned gmask = (DENSE_LANES == 32u) ? 0xFFFFFFFFu : ((1u << DENSE_LANES) - 1u);
    const int w = threadIdx.x >> 5;
    const unsigned slot = (unsigned)w * DENSE_GROUPS + grp;
    uint2* list = s_list[slot];
    uint2* sorted = s_sorted[slot];
    float* tw = &s_terms[slot][0][0];
    const float rscatter2 = P.rscatter * P.rscatter;
    const uint32_t stride = gridDim.x * 4u * DENSE_GROUPS;
    const uint32_t qn = *A.dense_n;
    publish_early<MULTI>(SL, qn);
    // The tail of the frame (finish_agent: walls, seek, integration, ~2/3 of an agent's instructions) is not done by one
    // lane per agent as the sums come out: the agents a warp works through are parked one per lane (slot index, state
    // and sums in registers) and finished together, every lane its own agent, when 32 are parked or the queue ends.
    uint32_t parked = 0;                       // agents parked in this warp (lane k holds the k-th)
    uint32_t my_p = 0, my_sp = 0;
    float4 my_me = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 my_key = make_uint2(0u, 0u);
    ForceSums my_acc{V(0.f, 0.f), V(0.f, 0.f), V(0.f, 0.f), 0.f};
    auto finish_parked = [&]() {
        if (lane < parked) {
            const uint32_t cell = finish_agent<KEEP_VEL, MULTI, false>(my_p, my_sp, my_me, my_key, A.vel4[my_sp], A.nav4[my_sp], my_acc, P, M, F, A, SL);
            A.cellrank[my_p] = make_uint2(cell, atomicAdd(F.count + cell, 1u));
        }
        parked = 0;
        __syncwarp();
    };
    for (uint32_t q_first = (blockIdx.x * 4u + (uint32_t)w) * DENSE_GROUPS; q_first < qn; q_first += stride) {
        const uint32_t qi = q_first + grp;
        const bool has = qi < qn;                   // the last warp of the queue may have an idle group
        const uint32_t p = has ? A.dense_q[qi] : A.dense_q[q_first];   // (an idle group shadows the first agent, writes nothing)
        const float4 me = A.pos4[p];
        const uint2 mykey = A.key2[p];
        const uint32_t sp = A.src[p];
        const v2 r = V(me.x, me.y);
        const int state_i = (int)(mykey.y & 3u);
        const int cxi = cs_cx(mykey.y), cyi = cs_cy(mykey.y);
        int fx, fy;
        fine_coords(F, r.x, r.y, fx, fy);
        const int xlo = max(fx - 1, 0), xhi = min(fx + 1, F.nx - 1);
        // ---- scan: the group's lanes stride over the three row ranges; accepted neighbours are appended with their keys
        uint32_t cnt = 0;
        for (int k = 0; k < 3; ++k) {
            const int yy = fy - 1 + k;
            uint32_t qb = 0u, qe = 0u;
            if (yy >= 0 && yy < F.ny) {
                const uint32_t rowbase = (uint32_t)yy * (uint32_t)F.nx;
                qb = __ldg(F.start + rowbase + xlo);
                qe = __ldg(F.start + rowbase + xhi + 1);
            }
            // warp-uniform trip count: the longer of the two groups' ranges
            uint32_t len = qe - qb;
            if (DENSE_GROUPS > 1u) len = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
            for (uint32_t off = 0; off < len; off += DENSE_LANES) {
                const uint32_t q = qb + off + gl;
                bool accept = false;
                uint32_t key = 0;
                if (q < qe) {
                    const float4 o = __ldg(A.pos4 + q);
                    const v2 dr = V(o.x, o.y) - r;
                    if (norm2(dr) < P.r_max2 && q != p) {
                        const uint2 ok = __ldg(A.key2 + q);
                        accept = cell_is_adjacent(ok.y, cxi, cyi);   // (only an agent outside the map can fail this, see k_step)
                        const int idx = (cs_cx(ok.y) - cxi + 1) * 3 + (cs_cy(ok.y) - cyi + 1);
                        const uint32_t rank = (idx < 4) ? (uint32_t)idx : (idx == 4 ? 8u : (uint32_t)(idx - 1));
                        key = (rank << 28) | ok.x;
                    }
                }
                const unsigned hits = (__ballot_sync(0xffffffffu, accept) >> gshift) & gmask;
                const uint32_t pos = cnt + (uint32_t)__popc(hits & ((1u << gl) - 1u));
                if (accept && pos < (uint32_t)DENSE_MAX) list[pos] = make_uint2(key, q);
                cnt += (uint32_t)__popc(hits);
            }
        }
        cnt = min(cnt, (uint32_t)DENSE_MAX);   // k_step queued this agent with the same count, <= DENSE_MAX
        uint32_t cnt_w = cnt;                  // warp-uniform bound for the loops below
        if (DENSE_GROUPS > 1u) cnt_w = max(cnt_w, __shfl_xor_sync(0xffffffffu, cnt_w, 16));
        // ---- rank sort into a second list: the keys are unique (gid), so an element's position is the number of smaller
        //      keys. Each lane takes two elements at a time and compares them with every key (one broadcast read each);
        //      no padding, no power-of-two rounding, no barrier per step — about half the instructions of the bitonic
        //      network for the 19..64 neighbours most crowded agents have.
        __syncwarp();
        for (uint32_t i0 = gl; i0 < cnt_w; i0 += 2u * DENSE_LANES) {
            const uint32_t i1 = i0 + DENSE_LANES;
            const uint2 e0 = (i0 < cnt) ? list[i0] : make_uint2(0xFFFFFFFFu, 0u);
            const uint2 e1 = (i1 < cnt) ? list[i1] : make_uint2(0xFFFFFFFFu, 0u);
            uint32_t r0 = 0u, r1 = 0u;
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint32_t kj = list[j].x;
                r0 += (kj < e0.x) ? 1u : 0u;
                r1 += (kj < e1.x) ? 1u : 0u;
            }
            if (i0 < cnt) sorted[r0] = e0;
            if (i1 < cnt) sorted[r1] = e1;
        }
        __syncwarp();
        // ---- force terms DENSE_LANES neighbours at a time (one per lane), then added in key order. Only the additions have
        //      to be sequential; the seven running sums (repulsion.xy, push.xy, dr_nn.xy, neighbour count) are independent
        //      of one another, so lanes 0..6 of the group each carry one of them through the chunk.
        float run = 0.f;                                                 // this lane's running sum
        for (uint32_t base = 0; base < cnt_w; base += DENSE_LANES) {
            const uint32_t k = base + gl;
            ForceTerms t;
            t.mask = 0u; t.rep = V(0.f, 0.f); t.push = V(0.f, 0.f); t.dr = V(0.f, 0.f);
            if (k < cnt) {
                const uint32_t q = sorted[k].y;
                t = neighbour_terms(r, fabsf(me.z), fabsf(me.w), state_i, __ldg(A.pos4 + q), rscatter2);
            }
            // A term the reference does not add is stored as +0: the running sums start at +0 and can never become -0
            // (x + (-x) and (+0) + (-0) round to +0), so adding +0 leaves every bit alone and the loop needs no test.
            const bool sc = (t.mask & 4u) != 0u;
            tw[0 * DENSE_LANES + gl] = t.rep.x; tw[1 * DENSE_LANES + gl] = t.rep.y;      // zero unless mask bit 0 (neighbour_terms)
            tw[2 * DENSE_LANES + gl] = t.push.x; tw[3 * DENSE_LANES + gl] = t.push.y;    // zero unless mask bit 1
            tw[4 * DENSE_LANES + gl] = sc ? t.dr.x : 0.f; tw[5 * DENSE_LANES + gl] = sc ? t.dr.y : 0.f;
            tw[6 * DENSE_LANES + gl] = sc ? 1.f : 0.f;
            __syncwarp();
            if (gl < 7u && base < cnt) {
                const uint32_t m = min(DENSE_LANES, cnt - base);
                for (uint32_t j = 0; j < m; ++j) run += tw[gl * DENSE_LANES + j];
            }
            __syncwarp();
        }
        ForceSums acc;
        acc.repulsion = V(__shfl_sync(0xffffffffu, run, gshift + 0), __shfl_sync(0xffffffffu, run, gshift + 1));
        acc.push = V(__shfl_sync(0xffffffffu, run, gshift + 2), __shfl_sync(0xffffffffu, run, gshift + 3));
        acc.dr_nn = V(__shfl_sync(0xffffffffu, run, gshift + 4), __shfl_sync(0xffffffffu, run, gshift + 5));
        acc.n_neighbours = __shfl_sync(0xffffffffu, run, gshift + 6);
        // park the agent(s) of this round: group g's agent goes to lane parked + g (every lane of a group holds its sums)
        for (uint32_t g = 0; g < DENSE_GROUPS; ++g) {
            const unsigned src_lane = g * DENSE_LANES;
            const bool g_has = __shfl_sync(0xffffffffu, has ? 1 : 0, src_lane) != 0;
            const uint32_t g_p = __shfl_sync(0xffffffffu, p, src_lane), g_sp = __shfl_sync(0xffffffffu, sp, src_lane);
            const float4 g_me = make_float4(__shfl_sync(0xffffffffu, me.x, src_lane), __shfl_sync(0xffffffffu, me.y, src_lane),
                                            __shfl_sync(0xffffffffu, me.z, src_lane), __shfl_sync(0xffffffffu, me.w, src_lane));
            const uint2 g_key = make_uint2(__shfl_sync(0xffffffffu, mykey.x, src_lane), __shfl_sync(0xffffffffu, mykey.y, src_lane));
            ForceSums g_acc;
            g_acc.repulsion = V(__shfl_sync(0xffffffffu, acc.repulsion.x, src_lane), __shfl_sync(0xffffffffu, acc.repulsion.y, src_lane));
            g_acc.push = V(__shfl_sync(0xffffffffu, acc.push.x, src_lane), __shfl_sync(0xffffffffu, acc.push.y, src_lane));
            g_acc.dr_nn = V(__shfl_sync(0xffffffffu, acc.dr_nn.x, src_lane), __shfl_sync(0xffffffffu, acc.dr_nn.y, src_lane));
            g_acc.n_neighbours = __shfl_sync(0xffffffffu, acc.n_neighbours, src_lane);
            if (g_has) {
                if (lane == parked) { my_p = g_p; my_sp = g_sp; my_me = g_me; my_key = g_key; my_acc = g_acc; }
                ++parked;
            }
        }
        if (parked + DENSE_GROUPS > 32u) finish_parked();
    }
    if (parked) finish_parked();
}

// Tolerance-mode counterpart of k_step_dense: FAST_DENSE_LANES lanes share the candidates of one crowded agent (every lane
// takes every FAST_DENSE_LANES-th candidate, branch-free terms as in k_step), the seven sums are reduced with shuffles —
// no lists, no ordering, no shared memory. The frame tail is run for 32 parked agents at a time, one per lane.
template <bool KEEP_VEL, bool MULTI>
__global__ void __launch_bounds__(128) k_step_dense_fast(RtsParams P, MapDev M, FineGrid F, AgentArrays A, SlabDev SL) {
    constexpr uint32_t GL = FAST_DENSE_LANES, GROUPS = 32u / GL;
    pdl_wait();
    pdl_trigger();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = lane / GL, gl = lane % GL;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t stride = gridDim.x * (blockDim.x >> 5) * GROUPS;
    const uint32_t qn = *A.dense_n;
    publish_early<MULTI>(SL, qn);
    const float rscatter2 = P.rscatter * P.rscatter;
    uint32_t parked = 0;                       // agents parked in this warp (lane k holds the k-th)
    uint32_t my_p = 0, my_sp = 0;
    float4 my_me = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 my_key = make_uint2(0u, 0u);
    ForceSums my_acc{V(0.f, 0.f), V(0.f, 0.f), V(0.f, 0.f), 0.f};
    auto finish_parked = [&]() {
        if (lane < parked) {
            const uint32_t cell = finish_agent<KEEP_VEL, MULTI, true>(my_p, my_sp, my_me, my_key, A.vel4[my_sp], A.nav4[my_sp], my_acc, P, M, F, A, SL);
            A.cellrank[my_p] = make_uint2(cell, atomicAdd(F.count + cell, 1u));
        }
        parked = 0;
        __syncwarp();
    };
    for (uint32_t q_first = warp * GROUPS; q_first < qn; q_first += stride) {
        const uint32_t qi = q_first + grp;
        const bool has = qi < qn;                   // the last warp of the queue may have idle groups
        const uint32_t p = A.dense_q[has ? qi : q_first];
        const float4 me = A.pos4[p];
        const uint2 mykey = A.key2[p];
        const uint32_t sp = A.src[p];
        int fx, fy;
        fine_coords(F, me.x, me.y, fx, fy);
        const int xlo = max(fx - 1, 0), xhi = min(fx + 1, F.nx - 1);
        FastScan c;
        c.r = V(me.x, me.y); c.radius = fabsf(me.z); c.mass_i = fabsf(me.w); c.standing = (mykey.y & 3u) == RTS_STANDING;
        c.r_max2 = P.r_max2; c.rscatter2 = rscatter2; c.self = p;
        c.near_edge = fx <= 1 || fx >= F.nx - 2 || fy <= 1 || fy >= F.ny - 2;
        c.cxi = cs_cx(mykey.y); c.cyi = cs_cy(mykey.y);
        uint32_t qb[3], ql[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int yy = fy - 1 + k;
            qb[k] = 0u; ql[k] = 0u;
            if (yy >= 0 && yy < F.ny) {
                const uint32_t rowbase = (uint32_t)yy * (uint32_t)F.nx;
                qb[k] = __ldg(F.start + rowbase + xlo);
                ql[k] = __ldg(F.start + rowbase + xhi + 1) - qb[k];
            }
        }
        c.qb0 = qb[0]; c.len0 = ql[0]; c.qb1 = qb[1]; c.len1 = ql[1]; c.qb2 = qb[2]; c.len2 = ql[2];
        ForceSums S{V(0.f, 0.f), V(0.f, 0.f), V(0.f, 0.f), 0.f};
        const unsigned lanes = __ballot_sync(0xffffffffu, has);
        if (has) fast_accumulate<false>(S, c, A.pos4, A.key2, gl, GL, lanes);
        __syncwarp();
#pragma unroll
        for (uint32_t off = GL >> 1; off > 0; off >>= 1) {   // sum over the lanes of the group
            S.repulsion.x += __shfl_xor_sync(0xffffffffu, S.repulsion.x, off); S.repulsion.y += __shfl_xor_sync(0xffffffffu, S.repulsion.y, off);
            S.push.x += __shfl_xor_sync(0xffffffffu, S.push.x, off); S.push.y += __shfl_xor_sync(0xffffffffu, S.push.y, off);
            S.dr_nn.x += __shfl_xor_sync(0xffffffffu, S.dr_nn.x, off); S.dr_nn.y += __shfl_xor_sync(0xffffffffu, S.dr_nn.y, off);
            S.n_neighbours += __shfl_xor_sync(0xffffffffu, S.n_neighbours, off);
        }
        // park the agents of this round: group g's agent goes to lane parked + g (every lane of a group holds its sums)
#pragma unroll
        for (uint32_t g = 0; g < GROUPS; ++g) {
            const unsigned src_lane = g * GL;
            const bool g_has = __shfl_sync(0xffffffffu, has ? 1 : 0, src_lane) != 0;
            const uint32_t g_p = __shfl_sync(0xffffffffu, p, src_lane), g_sp = __shfl_sync(0xffffffffu, sp, src_lane);
            const float4 g_me = make_float4(__shfl_sync(0xffffffffu, me.x, src_lane), __shfl_sync(0xffffffffu, me.y, src_lane),
                                            __shfl_sync(0xffffffffu, me.z, src_lane), __shfl_sync(0xffffffffu, me.w, src_lane));
            const uint2 g_key = make_uint2(__shfl_sync(0xffffffffu, mykey.x, src_lane), __shfl_sync(0xffffffffu, mykey.y, src_lane));
            ForceSums g_acc;
            g_acc.repulsion = V(__shfl_sync(0xffffffffu, S.repulsion.x, src_lane), __shfl_sync(0xffffffffu, S.repulsion.y, src_lane));
            g_acc.push = V(__shfl_sync(0xffffffffu, S.push.x, src_lane), __shfl_sync(0xffffffffu, S.push.y, src_lane));
            g_acc.dr_nn = V(__shfl_sync(0xffffffffu, S.dr_nn.x, src_lane), __shfl_sync(0xffffffffu, S.dr_nn.y, src_lane));
            g_acc.n_neighbours = __shfl_sync(0xffffffffu, S.n_neighbours, src_lane);
            if (g_has) {
                if (lane == parked) { my_p = g_p; my_sp = g_sp; my_me = g_me; my_key = g_key; my_acc = g_acc; }
                ++parked;
            }
        }
        if (parked + GROUPS > 32u) finish_parked();
    }
    if (parked) finish_parked();
}

// ------------------------------------------------------------------------------------------------
// rts_download: slot order -> caller order (dest = gid when gids are component indices, else slot)
struct Staging {
    float2* r; float2* vel; float2* inertia; float* angle; float* angle_vel; float* radius; float* mass;
    uint8_t* state; uint8_t* player; uint8_t* unit_type; int32_t* path_id; int32_t* waypoint; float2* r_next;
    uint32_t* gid; uint32_t* tri_idx; uint32_t* ns_cell; uint32_t* map_cell; uint32_t* flags;
};

__global__ void k_gather(uint32_t n, RtsParams P, AgentArrays A, Staging S, int by_gid, int clear_replan) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const float4 pos = A.pos4[d];
    const uint2 key = A.key2[d];
    const uint32_t sp = A.src[d];
    const float4 vel4 = A.vel4[sp];
    const float4 nav4 = A.nav4[sp];
    const uint32_t misc = __float_as_uint(nav4.w);
    const uint32_t o = by_gid ? key.x : d;
    if ((key.y >> 2) & 1u) {   // ghost (halo copy owned by the neighbour slab): not part of this handle's population
        if (S.gid) S.gid[o] = 0xFFFFFFFFu;
        return;
    }
    if (S.r) S.r[o] = make_float2(pos.x, pos.y);
    if (S.vel) S.vel[o] = A.vel_out ? A.vel_out[sp] : make_float2(0.f, 0.f);
    if (S.inertia) S.inertia[o] = make_float2(vel4.x, vel4.y);
    if (S.angle) S.angle[o] = vel4.z;
    if (S.angle_vel) S.angle_vel[o] = vel4.w;
    if (S.radius) S.radius[o] = fabsf(pos.z);   // (the sign bits mirror the state, see pack_pos)
    if (S.mass) S.mass[o] = fabsf(pos.w);
    if (S.state) S.state[o] = (uint8_t)(key.y & 3u);
    if (S.player) S.player[o] = (uint8_t)((misc >> 24) & 0xFu);
    if (S.unit_type) S.unit_type[o] = (uint8_t)((misc >> 16) & 0xFFu);
    if (S.path_id) S.path_id[o] = __float_as_int(nav4.z);
    if (S.waypoint) S.waypoint[o] = (int32_t)(misc & 0xFFFFu);
    if (S.r_next) S.r_next[o] = make_float2(nav4.x, nav4.y);
    if (S.gid) S.gid[o] = key.x;
    if (S.tri_idx) S.tri_idx[o] = A.tri[sp];
    if (S.ns_cell) S.ns_cell[o] = (uint32_t)cs_cx(key.y) + (uint32_t)cs_cy(key.y) * (uint32_t)P.ns_nx;
    if (S.map_cell) S.map_cell[o] = coord_to_cell(pos.x, pos.y, P.map_nx, P.map_ny, P.map_cell_size);
    if (S.flags) {
        const bool walked = P.walk_all || (key.y & 3u) == RTS_MOVING;
        S.flags[o] = ((misc >> 28) & 0xFu) | ((walked && A.tri[sp] == 0xFFFFFFFFu && A.n_tris_dev_hint) ? (uint32_t)RTS_FLAG_WALK_FAILED : 0u);
        if (clear_replan) A.nav4[sp].w = __uint_as_float(misc & ~((uint32_t)RTS_FLAG_REPLAN << 28));
    }
}

// rts_set_agents: caller-order staging -> the "n" buffers (index = caller index), then k_bin/scan/reorder
// rts_download_delta: the velocity of every agent in caller order, and {index, angle_vel, target, state} of the agents for which
// one of the three differs bitwise from `shadow` (what the host was last told; caller order), appended with a warp-aggregated
// atomic; `all` = report everyone (no valid shadow yet). The shadow is brought up to date on the way.
struct DeltaRec { uint32_t index; float angle_vel, tx, ty; uint32_t state; };
__global__ void __launch_bounds__(256) k_gather_delta(uint32_t n, AgentArrays A, float2* __restrict__ vel_out, float4* __restrict__ shadow,
                                                      DeltaRec* __restrict__ changes, uint32_t* __restrict__ counter, uint32_t cap, int all) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    bool changed = false;
    DeltaRec rec{0u, 0.f, 0.f, 0.f, 0u};
    if (d < n) {
        const uint2 key = A.key2[d];
        const uint32_t sp = A.src[d];
        const uint32_t o = key.x;   // component-index gids
        vel_out[o] = A.vel_out ? A.vel_out[sp] : make_float2(0.f, 0.f);
        const float av = A.vel4[sp].w;
        const float4 nav4 = A.nav4[sp];
        const uint32_t st = key.y & 3u;
        const float4 now = make_float4(av, nav4.x, nav4.y, __uint_as_float(st));
        const float4 old = shadow[o];
        changed = all || __float_as_uint(old.x) != __float_as_uint(now.x) || __float_as_uint(old.y) != __float_as_uint(now.y) ||
                  __float_as_uint(old.z) != __float_as_uint(now.z) || __float_as_uint(old.w) != __float_as_uint(now.w);
        if (changed) shadow[o] = now;
        rec = DeltaRec{o, av, nav4.x, nav4.y, st};
    }
    const unsigned need = __ballot_sync(0xffffffffu, changed);
    if (need) {
        const unsigned lane = threadIdx.x & 31u;
        uint32_t base = 0u;
        if (lane == (unsigned)(__ffs(need) - 1)) base = atomicAdd(counter, (uint32_t)__popc(need));
        base = __shfl_sync(0xffffffffu, base, __ffs(need) - 1);
        const uint32_t k = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
        if (changed && k < cap) changes[k] = rec;
    }
}

__global__ void k_ingest(uint32_t n, uint32_t first, Staging S, AgentArrays A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t d = first + i;
    const float2 r = S.r[i];
    const uint32_t state = S.state ? S.state[i] : (uint32_t)RTS_STANDING;
    A.pos4n[d] = pack_pos(r.x, r.y, S.radius ? S.radius[i] : 3.f, S.mass ? S.mass[i] : 1.f, state);
    A.key2n[d] = make_uint2(S.gid ? S.gid[i] : d, pack_cs(0u, 0u, 0u, state));
    const float2 in = S.inertia ? S.inertia[i] : make_float2(0.f, 0.f);
    A.vel4n[d] = make_float4(in.x, in.y, S.angle ? S.angle[i] : 0.f, S.angle_vel ? S.angle_vel[i] : 0.f);
    const float qnan = __uint_as_float(0x7FC00000u);
    const float2 rn = S.r_next ? S.r_next[i] : make_float2(qnan, qnan);
    const int32_t pid = S.path_id ? S.path_id[i] : -1;
    const uint32_t wp = S.waypoint ? (uint32_t)S.waypoint[i] : 0u;
    A.nav4n[d] = make_float4(rn.x, rn.y, __int_as_float(pid),
                             __uint_as_float(pack_misc(wp, S.unit_type ? S.unit_type[i] : 0u, S.player ? S.player[i] : 0u, 0u)));
    A.trin[d] = 0xFFFFFFFFu;
    if (A.vel_outn) A.vel_outn[d] = make_float2(0.f, 0.f);
}

// rts_update_agents (System2::updateSharedData): overwrite selected fields of every agent from
// caller-order staging (index = gid); positions are re-binned afterwards by the host code.
__global__ void k_apply(uint32_t n, Staging S, AgentArrays A) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    float4 pos = A.pos4[d];
    uint2 key = A.key2[d];
    const uint32_t sp = A.src[d];
    const uint32_t g = key.x;
    float4 vel4 = A.vel4[sp];
    float4 nav4 = A.nav4[sp];
    uint32_t misc = __float_as_uint(nav4.w);
    if (S.r) { const float2 r = S.r[g]; pos.x = r.x; pos.y = r.y; }
    if (S.radius) pos.z = S.radius[g];
    if (S.mass) pos.w = S.mass[g];
    if (S.state) key.y = (key.y & ~3u) | (S.state[g] & 3u);
    pos = pack_pos(pos.x, pos.y, pos.z, pos.w, key.y & 3u);
    if (S.inertia) { const float2 v = S.inertia[g]; vel4.x = v.x; vel4.y = v.y; }
    if (S.angle) vel4.z = S.angle[g];
    if (S.angle_vel) vel4.w = S.angle_vel[g];
    if (S.r_next) { const float2 v = S.r_next[g]; nav4.x = v.x; nav4.y = v.y; }
    if (S.path_id) nav4.z = __int_as_float(S.path_id[g]);
    if (S.waypoint) misc = (misc & ~0xFFFFu) | ((uint32_t)S.waypoint[g] & 0xFFFFu);
    if (S.unit_type) misc = (misc & ~(0xFFu << 16)) | ((uint32_t)S.unit_type[g] << 16);
    if (S.player) misc = (misc & ~(0xFu << 24)) | (((uint32_t)S.player[g] & 0xFu) << 24);
    nav4.w = __uint_as_float(misc);
    // the re-bin pass reads the "n" buffers at index = slot
    A.pos4n[d] = pos;
    A.key2n[d] = key;
    A.vel4n[d] = vel4;
    A.nav4n[d] = nav4;
    A.trin[d] = A.tri[sp];
    if (A.vel_outn) A.vel_outn[d] = A.vel_out[sp];
}

// rts_update_agents_indexed: the host touched m agents (commands, edits) — only their fields travel.
// k_slot_of_gid inverts slot -> gid (component-index gids), k_apply_indexed overwrites the non-null fields of the m listed
// agents IN PLACE (current order); with positions among the fields the caller re-bins afterwards (TO_N: the write goes to
// the "n" buffers that a preceding k_apply has filled in slot order).
__global__ void k_slot_of_gid(uint32_t n, const uint2* __restrict__ key2, uint32_t* __restrict__ inv) {
    const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) inv[key2[d].x] = d;
}
template <bool TO_N>
__global__ void k_apply_indexed(uint32_t m, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ inv, Staging S, AgentArrays A) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t d = inv[idx[k]];
    const uint32_t sp = TO_N ? d : A.src[d];
    float4* pos4 = TO_N ? A.pos4n : A.pos4;
    uint2* key2 = TO_N ? A.key2n : A.key2;
    float4* vel4p = TO_N ? A.vel4n : A.vel4;
    float4* nav4p = TO_N ? A.nav4n : A.nav4;
    float4 pos = pos4[d];
    uint2 key = key2[d];
    float4 vel4 = vel4p[sp];
    float4 nav4 = nav4p[sp];
    uint32_t misc = __float_as_uint(nav4.w);
    pos.z = fabsf(pos.z); pos.w = fabsf(pos.w);
    if (S.r) { const float2 r = S.r[k]; pos.x = r.x; pos.y = r.y; }
    if (S.radius) pos.z = S.radius[k];
    if (S.mass) pos.w = S.mass[k];
    if (S.state) key.y = (key.y & ~3u) | (S.state[k] & 3u);
    if (S.inertia) { const float2 v = S.inertia[k]; vel4.x = v.x; vel4.y = v.y; }
    if (S.angle) vel4.z = S.angle[k];
    if (S.angle_vel) vel4.w = S.angle_vel[k];
    if (S.r_next) { const float2 v = S.r_next[k]; nav4.x = v.x; nav4.y = v.y; }
    if (S.path_id) nav4.z = __int_as_float(S.path_id[k]);
    if (S.waypoint) misc = (misc & ~0xFFFFu) | ((uint32_t)S.waypoint[k] & 0xFFFFu);
    if (S.unit_type) misc = (misc & ~(0xFFu << 16)) | ((uint32_t)S.unit_type[k] << 16);
    if (S.player) misc = (misc & ~(0xFu << 24)) | (((uint32_t)S.player[k] & 0xFu) << 24);
    nav4.w = __uint_as_float(misc);
    pos4[d] = pack_pos(pos.x, pos.y, pos.z, pos.w, key.y & 3u);
    key2[d] = key;
    vel4p[sp] = vel4;
    nav4p[sp] = nav4;
}

// ------------------------------------------------------------------------------------------------
// Triangulation::updateCellGrid (Triangulation.cpp:1829-1864) on the device. The reference visits the cache cells in a
// zig-zag and locates each cell centre with findTriangle(centre, from_last_found = true): a serial chain, every walk
// starting where the previous one ended. The chain only matters for centres that lie on a triangle edge or vertex:
// a centre strictly inside a triangle (or on the outer boundary) is contained in exactly one triangle, and every
// terminating walk — and the linear-scan fallback — ends there. So:
//   k_cellgrid_classify  one thread per cell: all triangles containing the centre (inclusive test) among the ones whose
//                        bounding box touches the cell (MapDev scan lists); one hit -> final answer, several -> ambiguous
//   k_cellgrid_resolve   one thread per maximal run of ambiguous cells in zig-zag order: the reference's own walk from
//                        the predecessor's triangle, cell after cell (runs are short: a wall lying on a centre line)
__device__ __forceinline__ void zigzag_cell(uint32_t z, int nx, int ny, int& i, int& j) {
    j = (int)(z / (uint32_t)nx);
    const int k = (int)(z % (uint32_t)nx);
    const bool asc = (j % 2 == 0) && (j < ny - 1);   // row pairs ascending / descending; a left-over last row descends
    i = asc ? k : nx - 1 - k;
}
__device__ __forceinline__ iv2 cell_centre(int i, int j, float cs) {
    const float cx = (float)i * cs + cs / 2.f, cy = (float)j * cs + cs / 2.f;   // Triangulation.cpp:1838
    return iv2{(int32_t)cx, (int32_t)cy};
}
// Both kernels work on the zig-zag positions [z_lo, z_hi): the whole grid for a full build, the band of cache rows around a
// local change for rts_update_map_region (first_hit / ambiguous are indexed from z_lo; a run of ambiguous cells that begins
// before z_lo starts from that predecessor's current entry).
template <bool WIDE>
__global__ void k_cellgrid_classify(RtsParams P, MapDev M, uint32_t* cell2tri, uint32_t* first_hit, uint8_t* ambiguous, uint32_t z_lo, uint32_t z_hi) {
    const uint32_t z = z_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (z >= z_hi) return;
    int i, j;
    zigzag_cell(z, P.tri_nx, P.tri_ny, i, j);
    const iv2 q = cell_centre(i, j, P.tri_cell_size);
    const uint32_t cell = (uint32_t)j * (uint32_t)P.tri_nx + (uint32_t)i;
    uint32_t first = TRI_NONE, hits = 0;
    auto test = [&](uint32_t t) {
        iv2 a, b, c;
        int4 w1;
        load_tri_verts(M, t, a, b, c, w1);
        if (in_triangle<WIDE>(q, a, b, c)) {
            first = min(first, t);
            ++hits;
        }
    };
    const uint2 se = __ldg(M.scan_se + cell);
    for (uint32_t k = se.x, e = se.y; k < e; ++k) test(__ldg(M.scan_list + k));
    for (uint32_t k = 0; k < M.n_scan_large; ++k) test(__ldg(M.scan_large + k));
    cell2tri[cell] = first;
    first_hit[z - z_lo] = first;
    ambiguous[z - z_lo] = hits > 1 ? 1 : 0;
}
template <bool WIDE>
__global__ void k_cellgrid_resolve(RtsParams P, MapDev M, uint32_t last_found, uint32_t* cell2tri, const uint32_t* first_hit,
                                   const uint8_t* ambiguous, uint32_t z_lo, uint32_t z_hi) {
    const uint32_t z0 = z_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (z0 >= z_hi || !ambiguous[z0 - z_lo] || (z0 > z_lo && ambiguous[z0 - z_lo - 1])) return;   // heads of ambiguous runs only
    int i, j;
    uint32_t prev = last_found;
    if (z0 > 0) {   // the predecessor is unambiguous (or outside the band), its entry is final
        zigzag_cell(z0 - 1, P.tri_nx, P.tri_ny, i, j);
        prev = cell2tri[(uint32_t)j * (uint32_t)P.tri_nx + (uint32_t)i];
    }
    if (prev >= M.n_tris) prev = TRI_NONE;
    for (uint32_t z = z0; z < z_hi && ambiguous[z - z_lo]; ++z) {
        zigzag_cell(z, P.tri_nx, P.tri_ny, i, j);
        const iv2 q = cell_centre(i, j, P.tri_cell_size);
        uint32_t flags = 0;
        uint32_t t = find_triangle_from<WIDE>(M, q, prev, flags);
        if (t == TRI_NEED_SCAN) t = first_hit[z - z_lo];   // the reference's linear scan returns the lowest containing index
        cell2tri[(uint32_t)j * (uint32_t)P.tri_nx + (uint32_t)i] = t;
        if (t != TRI_NONE) prev = t;
    }
}

// rts_update_map_region: a few records of a device table are replaced (index list + values staged by the host)
__global__ void k_scatter_u32(uint32_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ val, uint32_t* __restrict__ dst, int or_into) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = or_into ? (dst[idx[i]] | val[i]) : val[i];
}
__global__ void k_scatter_u2(uint32_t n, const uint32_t* __restrict__ idx, const uint2* __restrict__ val, uint2* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = val[i];
}
__global__ void k_scatter_tris(uint32_t n, const uint32_t* __restrict__ idx, const int4* __restrict__ val, int4* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t t = idx[i];
    dst[3 * t] = val[3 * i]; dst[3 * t + 1] = val[3 * i + 1]; dst[3 * t + 2] = val[3 * i + 2];
}

// ---- f4, first slice: the two neighbour queries next to the frame path, on the cell order the frame maintains anyway ----
// NeighbourSearcherContext::getNeighboursIndsFull (NeighbourSearcherContext.h:44-92) on the physics grid: one thread per
// query scans the fine cells its disk touches and emits {physics cell index, gid} of every agent with dist2 <= r_max^2 whose
// physics cell lies in the block of cells the reference searches; the host orders each result list by that pair, which is
// the reference's visit order (cells row-major, ascending component index inside a cell).
__global__ void k_query_radius(uint32_t nq, const float2* __restrict__ q_xy, const float* __restrict__ q_r, uint32_t cap, RtsParams P, FineGrid F,
                               AgentArrays A, uint32_t* __restrict__ counts, uint2* __restrict__ out) {
    const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const v2 q = V(q_xy[qi].x, q_xy[qi].y);
    const float r_max = q_r[qi], r_max2 = r_max * r_max;
    const uint32_t cell_ind = coord_to_cell(q.x, q.y, P.ns_nx, P.ns_ny, P.ns_cell_size);
    const int ccx = (int)(cell_ind % (uint32_t)P.ns_nx), ccy = (int)(cell_ind / (uint32_t)P.ns_nx);
    int j_max = ccy + (int)ceilf(r_max / P.ns_cell_size) + 1, j_min = ccy - (int)ceilf(r_max / P.ns_cell_size) - 1;
    j_max = min(j_max, P.ns_ny - 1);
    j_min = max(j_min, 0);
    const int delta = j_max - j_min;
    const int i_max = min(ccx + delta, P.ns_nx - 1), i_min = max(ccx - delta, 0);
    int fx0, fy0, fx1, fy1;
    fine_coords(F, q.x - r_max, q.y - r_max, fx0, fy0);
    fine_coords(F, q.x + r_max, q.y + r_max, fx1, fy1);
    uint32_t cnt = 0;
    for (int fy = fy0; fy <= fy1; ++fy) {
        const uint32_t rowbase = (uint32_t)fy * (uint32_t)F.nx;
        for (uint32_t s = __ldg(F.start + rowbase + fx0), e = __ldg(F.start + rowbase + fx1 + 1); s < e; ++s) {
            const float4 o = A.pos4[s];
            if (!(dist2(V(o.x, o.y), q) <= r_max2)) continue;
            const uint2 key = A.key2[s];
            if ((key.y >> 2) & 1u) continue;   // ghost
            const int cx = cs_cx(key.y), cy = cs_cy(key.y);
            if (cx < i_min || cx > i_max || cy < j_min || cy > j_max) continue;
            if (cnt < cap) out[(size_t)qi * cap + cnt] = make_uint2((uint32_t)cx + (uint32_t)cy * (uint32_t)P.ns_nx, key.x);
            ++cnt;
        }
    }
    counts[qi] = cnt;
}
// NeighbourSearcherStrategyAttack::execute (NeighbourSearcherStrategy.cpp:125-177) as written: the enemy its loop ends on =
// the agent of another player with dist2 < range2 that comes LAST in visit order (calcNearestCells cells, own cell last,
// ascending component index). One thread per agent; the player of a candidate is fetched only when its key beats the best.
__global__ void k_attack_targets(uint32_t n, float range2, RtsParams P, FineGrid F, AgentArrays A, int32_t* __restrict__ out, int by_gid) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float4 me = A.pos4[p];
    const uint2 mykey = A.key2[p];
    if ((mykey.y >> 2) & 1u) return;
    const uint32_t my_player = (__float_as_uint(A.nav4[A.src[p]].w) >> 24) & 0xFu;
    const v2 r = V(me.x, me.y);
    const int cxi = cs_cx(mykey.y), cyi = cs_cy(mykey.y);
    const float reach = sqrtf(range2);
    int fx0, fy0, fx1, fy1;
    fine_coords(F, r.x - reach, r.y - reach, fx0, fy0);
    fine_coords(F, r.x + reach, r.y + reach, fx1, fy1);
    uint32_t best_key = 0u;
    int32_t best = -1;
    for (int fy = fy0; fy <= fy1; ++fy) {
        const uint32_t rowbase = (uint32_t)fy * (uint32_t)F.nx;
        for (uint32_t s = __ldg(F.start + rowbase + fx0), e = __ldg(F.start + rowbase + fx1 + 1); s < e; ++s) {
            const float4 o = A.pos4[s];
            if (!(dist2(V(o.x, o.y), r) < range2)) continue;
            const uint2 key = A.key2[s];
            if (((key.y >> 2) & 1u) || !cell_is_adjacent(key.y, cxi, cyi)) continue;
            const int idx = (cs_cx(key.y) - cxi + 1) * 3 + (cs_cy(key.y) - cyi + 1);
            const uint32_t rank = (idx < 4) ? (uint32_t)idx : (idx == 4 ? 8u : (uint32_t)(idx - 1));
            const uint32_t k = ((rank << 28) | key.x) + 1u;   // (+1: a key of 0 means "none yet")
            if (k <= best_key) continue;
            const uint32_t pl = (__float_as_uint(A.nav4[A.src[s]].w) >> 24) & 0xFu;
            if (pl == my_player) continue;
            best_key = k;
            best = (int32_t)key.x;
        }
    }
    out[by_gid ? mykey.x : p] = best;
}

// stand-alone queries
// MapDev::border_top / border_right: the reference's walk (no shortcut through the table being filled) for every integer point
// on the two far border lines of the cache grid
template <bool WIDE>
__global__ void __launch_bounds__(128) k_border_cache(RtsParams P, MapDev M, uint32_t* top, uint32_t* right, int32_t gw, int32_t gh) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_top = (uint32_t)gw + 1u, total = n_top + (uint32_t)gh + 1u;
    const bool want = i < total;
    const float fx = (i < n_top) ? (float)i : (float)gw;
    const float fy = (i < n_top) ? (float)gh : (float)(i - n_top);
    uint32_t flags = 0u;
    const uint32_t t = find_triangle_warp<WIDE, false>(M, P, want, fx, fy, flags, 0xffffffffu);
    if (want) {
        if (i < n_top) top[i] = t;
        else right[i - n_top] = t;
    }
}

__global__ void k_coord_to_cell(const float2* xy, uint32_t n, int32_t nx, int32_t ny, float cs, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = coord_to_cell(xy[i].x, xy[i].y, nx, ny, cs);
}
__global__ void k_find_triangles(const float2* xy, uint32_t n, RtsParams P, MapDev M, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned mask = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    uint32_t flags = 0;
    const float2 q = xy[i];
    const bool small = M.coords_fit_i32 && fabsf(q.x) <= 16383.f && fabsf(q.y) <= 16383.f;
    if (__all_sync(mask, small)) out[i] = find_triangle_warp<false>(M, P, true, q.x, q.y, flags, mask);
    else out[i] = find_triangle_warp<true>(M, P, true, q.x, q.y, flags, mask);
}
__global__ void k_contact_edges(const float2* xy, const float* radius, uint32_t n, uint32_t cap, RtsParams P, MapDev M,
                                uint32_t* counts, float* edges_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = 0;
    float* o = edges_out ? edges_out + (size_t)5 * cap * i : nullptr;
    counts[i] = (uint32_t)contact_walls(M, P, V(xy[i].x, xy[i].y), radius[i], [&](float, float, uint32_t e) {
        if (o && k < cap) {
            const float4 ft = M.edge_ft[e];
            o[5 * k + 0] = ft.x; o[5 * k + 1] = ft.y; o[5 * k + 2] = ft.z; o[5 * k + 3] = ft.w; o[5 * k + 4] = M.edge_l[e];
        }
        ++k;
    });
}

}  // namespace rts
