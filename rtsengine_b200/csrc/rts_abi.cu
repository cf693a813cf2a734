// C ABI of librts_b200.so (include/rts_gpu.h): context, table uploads, step scheduling, downloads.
// Host code is plain C++; nothing here falls back to the CPU — without a CUDA device every entry point
// that computes returns RTS_E_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is resolved with dlopen when rts_comm_init is called

#include "rts_kernels.cuh"

using namespace rts;

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return h->fail(RTS_E_CUDA, #call, cudaGetErrorString(e_));          \
    } while (0)

namespace {

struct KernelTimer {  // per-kind CUDA-event timing of launches on the handle's stream (profiling mode only)
    std::vector<cudaEvent_t> beg, end;
    size_t used = 0;
    double total_ms = 0;
    uint64_t launches = 0;
};

enum { KT_STEP = 0, KT_BIN = 1, KT_SCAN = 2, KT_REORDER = 3, KT_EXCHANGE = 4, KT_INGEST = 5, KT_WALK = 6, KT_DENSE = 7, KT_COUNT = 8 };

}  // namespace

struct RtsContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    RtsParams P{};
    MapDev M{};
    FineGrid F{};
    AgentArrays A{};
    uint32_t n = 0, cap = 0;
    uint32_t* c2t_dev = nullptr;   // writable alias of M.cell2tri
    bool map_ready = false, gid_is_index = true, keep_vel = true, profiling = false;
    float fine_cs = 6.0f;
    bool fine_cs_auto = true;         // no explicit RtsParams.reserved[1]: the gather-cell size follows the population density (rts_set_agents)
    std::string err;
    // owned device allocations of the tables
    std::vector<void*> table_allocs, path_allocs, type_allocs;
    // rts_download_delta: what the host was last told (caller order), the change list, its counter (device + pinned host copy)
    float4* delta_shadow = nullptr; DeltaRec* delta_changes = nullptr; uint32_t* delta_count = nullptr; uint32_t* delta_count_host = nullptr;
    uint32_t delta_cap = 0; bool delta_valid = false; cudaEvent_t ev_delta = nullptr;
    uint32_t* border_dev = nullptr; size_t border_cap = 0;     // MapDev::border_top / border_right (one allocation)
    void* path_arena = nullptr; size_t path_arena_cap = 0;      // rts_upload_paths: the path tables, one allocation
    void* group_arena = nullptr; uint32_t group_arena_ng = 0;   // ... and the groups' arrival state (kept across re-uploads)
    std::vector<char> path_image;                                // host image of the path arena
    // host mirrors + writable device aliases of the map tables: rts_update_map_region patches them in place
    std::vector<RtsTriangle> tris_host;
    std::vector<uint2> line_se_host, scan_se_host;
    std::vector<uint32_t> wall_bits_host;
    uint32_t tri_cap = 0, edge_pool_cap = 0, edge_pool_used = 0, scan_pool_cap = 0, scan_pool_used = 0, scan_large_cap = 0, scan_outer_cap = 0;
    int4* tris_w = nullptr; uint2* line_se_w = nullptr; float4* edge_ft_w = nullptr; float* edge_l_w = nullptr;
    uint2* scan_se_w = nullptr; uint32_t* scan_list_w = nullptr; uint32_t* scan_large_w = nullptr; uint32_t* scan_outer_w = nullptr;
    uint32_t* wall_bits_w = nullptr;
    uint64_t upload_bytes_last = 0, upload_bytes_full = 0;   // H2D bytes of the last map call / of the last full rts_upload_map
    std::vector<void*> agent_allocs;
    // staging for upload / download (caller order), sized to cap
    Staging S{};
    std::vector<void*> staging_allocs;
    uint32_t staging_cap = 0;
    void* scratch = nullptr;  // generic device scratch for the stand-alone queries
    size_t scratch_bytes = 0;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t stream2 = nullptr, stream4 = nullptr;   // (unused since the walk moved into the frame kernel) / slab edge rows + exchange
    cudaEvent_t ev_scatter = nullptr, ev_walk = nullptr, ev_fork2 = nullptr, ev_comm = nullptr;
    std::vector<RtsUnitType> types_host;
    float max_type_speed = 100.f;     // max over unit types of the velocity clamp + seek speed, see set_edge_rows
    bool ingest_done = false;         // k_ingest_remote of this frame has been launched
    bool graph_slabs = false;         // reserved[5] bit 4 (value 16): slab frames, NCCL calls included, replayed from a CUDA graph too
    bool frame_is_split = false;      // this frame's edge rows, exchange and ingest are on stream4 (see split_frame)
    bool overlap_exchange = false;    // reserved[5] bit 3 (value 8) turns the split on: at 2 GPUs it measured the same as exchanging
                                      // between the frame kernel and the scan (0.2503 vs 0.2511 ms per frame), so it is opt-in
    bool walk_pending = false;        // ev_walk has been recorded and not yet waited for by the main stream
    bool walkq_pending = false;       // the walk queue of the last frame has not been searched yet (the next frame kernel will,
    int walkq_parity = 0;             // ... or join_walk before anything reads the triangle indices)
    KernelTimer kt[KT_COUNT];
    double last_step_ms = 0;
    uint64_t launches = 0, steps = 0;
    // step graph (two frames: the previous-order buffers ping-pong)
    cudaGraphExec_t graph2[2] = {nullptr, nullptr};   // GRAPH_FRAMES frames, one per ping-pong parity
    uint32_t graph_n[2] = {0, 0};
    uint64_t graph_launches = 0;
    cudaGraphExec_t graph1[2] = {nullptr, nullptr};   // ONE frame (rts_step(h, 1), the per-frame call of the System2 adapter), one per parity
    uint32_t graph1_n[2] = {0, 0};
    uint64_t graph1_launches = 0;
    int parity = 0;
    bool use_graph = true;
    // ---- slab decomposition (multi-GPU) ----
    bool multi = false;
    bool overlap_walk = true;         // walk queue searched by the next frame kernel's leading CTAs (reserved[4] = 1: k_walk after k_scatter)
    bool fast = false;                // RtsParams.precision == RTS_PRECISION_FAST: tolerance-mode kernels
    bool pdl = true;                  // frame kernels launched with programmatic stream serialization (reserved[5] bit 6, value 64, turns it off)
    int walk_ctas_per_sm = 4;         // (RTS_WALK_CTAS_PER_SM overrides: tuning runs)
    int n_sm = 148;                   // cudaDevAttrMultiProcessorCount of the handle's device (grids of the persistent-style kernels)
    SlabDev SL{};
    uint32_t cap_main = 0;            // slots the frame kernel may use; arrivals are appended behind them
    std::vector<void*> slab_allocs;
    size_t ex_bytes = 0;              // bytes of one exchange buffer (header + migrant + halo segments)
    void* ex_send[2] = {nullptr, nullptr};
    void* ex_recv[2] = {nullptr, nullptr};
    uint32_t* n_dev = nullptr;
    // peer exchange over NVLink (CUDA IPC): see SlabDev
    void* arena = nullptr;            // this rank's receive buffers (2 directions x 2 parities) + flags + epoch words, exported to the neighbours
    void* peer_arena[2] = {nullptr, nullptr};   // the neighbours' arenas as mapped into this process
    bool want_peer = true;            // reserved[5] bit 5 (value 32): keep the per-frame ncclSend/ncclRecv exchange instead
    bool exchange_pending = false;    // a peer exchange has been launched since the last k_scatter (which ticks the epoch)
    // NCCL, resolved at run time
    void* nccl_lib = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    decltype(&ncclCommInitRank) p_ncclCommInitRank = nullptr;
    decltype(&ncclCommDestroy) p_ncclCommDestroy = nullptr;
    decltype(&ncclSend) p_ncclSend = nullptr;
    decltype(&ncclRecv) p_ncclRecv = nullptr;
    decltype(&ncclGroupStart) p_ncclGroupStart = nullptr;
    decltype(&ncclGroupEnd) p_ncclGroupEnd = nullptr;
    decltype(&ncclGetErrorString) p_ncclGetErrorString = nullptr;
    decltype(&ncclAllGather) p_ncclAllGather = nullptr;
    decltype(&ncclAllReduce) p_ncclAllReduce = nullptr;

    void drop_graphs() {
        for (auto& g : graph2)
            if (g) { cudaGraphExecDestroy(g); g = nullptr; }
        for (auto& g : graph1)
            if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    }

    int32_t fail(int32_t code, const char* what, const char* detail) {
        err = std::string(what) + ": " + detail;
        return code;
    }
};

namespace {

template <class T>
int32_t dev_alloc(RtsContext* h, T** p, size_t count, std::vector<void*>& owner) {
    void* q = nullptr;
    CU(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    owner.push_back(q);
    *p = static_cast<T*>(q);
    return RTS_OK;
}

void free_all(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

int32_t ensure_scratch(RtsContext* h, size_t bytes) {
    if (bytes <= h->scratch_bytes) return RTS_OK;
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr;
    h->scratch_bytes = 0;
    CU(cudaMalloc(&h->scratch, bytes));
    h->scratch_bytes = bytes;
    return RTS_OK;
}

void timer_begin(RtsContext* h, int kind) {
    if (!h->profiling) return;
    KernelTimer& t = h->kt[kind];
    if (t.used == t.beg.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        t.beg.push_back(a);
        t.end.push_back(b);
    }
    cudaEventRecord(t.beg[t.used], h->stream);
}
void timer_end(RtsContext* h, int kind) {
    if (!h->profiling) return;
    KernelTimer& t = h->kt[kind];
    cudaEventRecord(t.end[t.used], h->stream);
    t.used++;
}
void timer_collect(RtsContext* h) {  // after a stream sync
    for (auto& t : h->kt) {
        for (size_t i = 0; i < t.used; ++i) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, t.beg[i], t.end[i]) == cudaSuccess) {
                t.total_ms += ms;
                t.launches++;
            }
        }
        t.used = 0;
    }
}

inline uint32_t blocks_for(uint32_t n, uint32_t b) { return (n + b - 1) / b; }

// Launch of a kernel of the frame chain (k_step -> k_step_dense* -> [k_ingest_remote] -> k_scan_onepass -> k_scatter -> k_step ...).
// With `pdl` the launch carries cudaLaunchAttributeProgrammaticStreamSerialization: the kernel's CTAs may become resident
// while the previous kernel of the stream drains its last CTAs (launch latency, block scheduling and the prologue overlap
// that tail); every such kernel executes griddepcontrol.wait (pdl_wait) before it touches global memory, which blocks until
// the previous kernel — and, transitively, everything before it — has completed and flushed. Works inside stream capture.
template <typename... KArgs, typename... Args>
inline void launch_chain(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
inline bool use_pdl(const RtsContext* h) { return h->pdl && !h->profiling; }
inline uint32_t walk_ctas_of(const RtsContext* h) { return (uint32_t)h->n_sm * (uint32_t)h->walk_ctas_per_sm; }   // CTAs at the head of k_step's grid that search the previous frame's walk queue

int32_t alloc_agents(RtsContext* h, uint32_t cap) {
    free_all(h->agent_allocs);
    h->cap = 0;
    AgentArrays& A = h->A;
    A = AgentArrays{};
    int32_t rc;
#define AL(field, T)                                                              \
    if ((rc = dev_alloc<T>(h, &A.field, cap, h->agent_allocs)) != RTS_OK) return rc;
    AL(pos4, float4) AL(key2, uint2) AL(src, uint32_t) AL(vel4, float4) AL(nav4, float4) AL(tri, uint32_t)
    AL(pos4n, float4) AL(key2n, uint2) AL(vel4n, float4) AL(nav4n, float4) AL(trin, uint32_t) AL(cellrank, uint2)
    if (h->keep_vel) { AL(vel_out, float2) AL(vel_outn, float2) }
    AL(stop_req, uint32_t)
#undef AL
    {   // work list of the triangle walk + its two alternating counters
        if ((rc = dev_alloc<uint32_t>(h, &A.walk_q, (size_t)cap, h->agent_allocs)) != RTS_OK) return rc;
        uint32_t* wn = nullptr;
        if ((rc = dev_alloc<uint32_t>(h, &wn, 4, h->agent_allocs)) != RTS_OK) return rc;
        CU(cudaMemsetAsync(wn, 0, 16, h->stream));
        A.walk_n = wn;
    }
    {   // two queues of crowded agents (whole population / interior, and the edge rows of a slab) + their counters
        if ((rc = dev_alloc<uint32_t>(h, &A.dense_q, 2 * (size_t)cap, h->agent_allocs)) != RTS_OK) return rc;
        uint32_t* dn = nullptr;
        if ((rc = dev_alloc<uint32_t>(h, &dn, 64, h->agent_allocs)) != RTS_OK) return rc;
        CU(cudaMemsetAsync(dn, 0, 256, h->stream));
        A.dense_n = dn;
        A.work = dn + 32;   // tile counter of the persistent k_step, 128 bytes away from the queue counters
    }
    CU(cudaMemsetAsync(A.stop_req, 0, sizeof(uint32_t) * (size_t)cap, h->stream));
    A.n_dev = h->n_dev;
    A.n_tris_dev_hint = h->map_ready ? h->M.n_tris : 0u;
    h->cap = cap;
    h->drop_graphs();
    return RTS_OK;
}

int32_t alloc_staging(RtsContext* h, uint32_t cap) {
    if (cap <= h->staging_cap) return RTS_OK;
    free_all(h->staging_allocs);
    h->staging_cap = 0;
    Staging& S = h->S;
    int32_t rc;
#define AL(field, T) \
    if ((rc = dev_alloc<T>(h, &S.field, cap, h->staging_allocs)) != RTS_OK) return rc;
    AL(r, float2) AL(vel, float2) AL(inertia, float2) AL(angle, float) AL(angle_vel, float) AL(radius, float) AL(mass, float)
    AL(state, uint8_t) AL(player, uint8_t) AL(unit_type, uint8_t) AL(path_id, int32_t) AL(waypoint, int32_t)
    AL(r_next, float2) AL(gid, uint32_t) AL(tri_idx, uint32_t) AL(ns_cell, uint32_t) AL(map_cell, uint32_t) AL(flags, uint32_t)
#undef AL
    h->staging_cap = cap;
    return RTS_OK;
}

int32_t setup_fine_grid(RtsContext* h) {
    const RtsParams& P = h->P;
    float cs = h->fine_cs;
    const float need = std::sqrt(P.r_max2) * 1.05f;
    if (cs < need) cs = need;
    FineGrid F{};
    F.inv_cs = 1.0f / cs;
    const float W = P.map_nx * P.map_cell_size, H = P.map_ny * P.map_cell_size;
    F.nx = (int32_t)(W / cs) + 2;
    // slab: physics rows [lo, hi) plus one physics row of halo on each side
    const float y_lo = std::max(0.f, (P.slab_row_lo - 1) * P.ns_cell_size);
    const float y_hi = std::min(H + cs, (P.slab_row_hi + 1) * P.ns_cell_size);
    F.row0 = (int32_t)(y_lo / cs);
    F.ny = (int32_t)(y_hi / cs) + 2 - F.row0;
    F.n_cells = (uint32_t)F.nx * (uint32_t)F.ny;
    if (h->F.count) { cudaFree(h->F.count); cudaFree(h->F.start); cudaFree(h->F.partial); cudaFree(h->F.status); cudaFree(h->F.ticket); }
    CU(cudaMalloc(&F.status, sizeof(unsigned long long) * (blocks_for(F.n_cells + 8, SCAN_TILE) + 1)));
    CU(cudaMalloc(&F.ticket, 16));
    CU(cudaMemsetAsync(F.status, 0, sizeof(unsigned long long) * (blocks_for(F.n_cells + 8, SCAN_TILE) + 1), h->stream));
    CU(cudaMemsetAsync(F.ticket, 0, 16, h->stream));
    const size_t n_pad = ((size_t)F.n_cells + 1 + 7) & ~(size_t)7;   // see k_scan_onepass
    CU(cudaMalloc(&F.count, sizeof(uint32_t) * n_pad));
    CU(cudaMalloc(&F.start, sizeof(uint32_t) * n_pad));
    CU(cudaMalloc(&F.partial, sizeof(uint32_t) * (blocks_for(F.n_cells, SCAN_TILE) + 1)));
    CU(cudaMemsetAsync(F.count, 0, sizeof(uint32_t) * n_pad, h->stream));
    CU(cudaMemsetAsync(F.start, 0, sizeof(uint32_t) * n_pad, h->stream));
    h->F = F;
    h->drop_graphs();
    return RTS_OK;
}

void swap_prev_order(RtsContext* h) {
    AgentArrays& A = h->A;
    std::swap(A.vel4, A.vel4n);
    std::swap(A.nav4, A.nav4n);
    std::swap(A.tri, A.trin);
    std::swap(A.vel_out, A.vel_outn);
    h->parity ^= 1;
}

void launch_walk_kernel(RtsContext* h, cudaStream_t st, int parity) {
    if (h->M.coords_fit_i32) k_walk<false><<<h->n_sm * 4, 256, 0, st>>>(h->P, h->M, h->A, parity);
    else k_walk<true><<<h->n_sm * 4, 256, 0, st>>>(h->P, h->M, h->A, parity);
    h->launches += 1;
}
// called before anything reads the triangle indices or overwrites pos4/src: searches a walk queue that is still pending
int32_t join_walk(RtsContext* h) {
    if (h->walk_pending) {   // (legacy: a walk kernel on stream2)
        CU(cudaStreamWaitEvent(h->stream, h->ev_walk, 0));
        h->walk_pending = false;
    }
    if (h->walkq_pending) {   // nobody has searched for the agents the last k_scatter queued: do it now
        launch_walk_kernel(h, h->stream, h->walkq_parity);
        h->walkq_pending = false;
    }
    return RTS_OK;
}

// scan of the counts + scatter into the new order; the "n" buffers become the previous-order arrays
int32_t launch_scan_reorder(RtsContext* h) {
    const uint32_t nb = blocks_for(((h->F.n_cells + 1u + 7u) & ~7u), SCAN_TILE);
    timer_begin(h, KT_SCAN);
    launch_chain(use_pdl(h), k_scan_onepass, dim3(nb), dim3(SCAN_THREADS), h->stream, h->F, h->n_dev);
    timer_end(h, KT_SCAN);
    int32_t rc = join_walk(h);   // a pending walk queue refers to what the scatter is about to overwrite (normally none is pending)
    if (rc != RTS_OK) return rc;
    timer_begin(h, KT_REORDER);
    const uint32_t n_slots = h->multi ? h->cap_main + 2 * (h->SL.cap_mig + h->SL.cap_halo) : h->n;
    const int wpar = h->parity;
    const int do_walk = h->map_ready ? 1 : 0;
    const int tick = (h->multi && h->SL.peer_mode && h->exchange_pending) ? 1 : 0;
    h->exchange_pending = false;
    if (n_slots) launch_chain(use_pdl(h), k_scatter, dim3(blocks_for(n_slots, 256)), dim3(256), h->stream, n_slots, h->F, h->A, h->SL, h->multi ? 1 : 0, h->P, h->M, wpar, do_walk, tick);
    timer_end(h, KT_REORDER);
    h->launches += 2;
    swap_prev_order(h);
    // Triangle walk of the new order: k_scatter has confirmed most agents' triangle from last frame's and queued the rest.
    // By default the queue is left to the leading CTAs of the NEXT frame kernel (launch_step_interior), so the walk
    // overlaps the frame inside one stream; join_walk searches it earlier if something needs the indices before that.
    // reserved[4] / profiling: a k_walk launch right here instead.
    if (n_slots && h->map_ready) {
        // (only with walk_all: every agent's carried index is re-verified by the next k_scatter, so it does not matter that
        //  the frame kernel may copy an entry forward before the search has replaced it)
        if (h->P.walk_all != 0 && h->overlap_walk && !h->profiling) {
            h->walkq_pending = true;
            h->walkq_parity = wpar;
        } else {
            timer_begin(h, KT_WALK);
            launch_walk_kernel(h, h->stream, wpar);
            timer_end(h, KT_WALK);
        }
    }
    return RTS_OK;
}

// Width of the band along a slab boundary whose agents are mirrored on the neighbour: the interaction range; with the arrival
// rule on, finish_radius — an agent can stop (and be stopped by) same-group agents that far away (k_arrive).
float halo_band(const RtsParams& P) {
    const float physics = std::sqrt(P.r_max2) * 1.05f + 0.25f;
    return P.enable_arrival ? std::max(physics, P.finish_radius + 0.25f) : physics;
}

// Edge rows of a slab (SlabDev): the fine rows that reach into the halo band plus the farthest an agent can travel in one
// frame (velocity clamp of the fastest unit type + its seek speed), with one more fine row of margin.
void set_edge_rows(RtsContext* h) {
    SlabDev& SL = h->SL;
    const FineGrid& F = h->F;
    const float max_travel = h->max_type_speed * std::fabs(h->P.dt);
    const float reach = SL.halo + max_travel + 1.0f / F.inv_cs;
    const int lo = (int)std::floor((SL.y_lo + reach) * F.inv_cs) - F.row0 + 1;
    const int hi = (int)std::floor((SL.y_hi - reach) * F.inv_cs) - F.row0;
    SL.edge_row_lo = std::min(std::max(lo, 1), F.ny - 1);
    SL.edge_row_hi = std::min(std::max(hi, SL.edge_row_lo), F.ny - 1);
    SL.interior_launch = 0;
    // grid of the edge launch, per side: the edge rows hold ~3x the agents of the halo band (band + travel + one row +
    // the ghosts below the boundary); the halo capacity already carries the caller's safety factor
    SL.edge_blocks = std::min(blocks_for(2u * SL.cap_halo + SL.cap_mig, STEP_BLOCK), blocks_for(std::max(h->cap_main, 1u), STEP_BLOCK));
}

// Does this frame split the frame kernel into the edge rows (comm stream, followed there by the exchange and the ingest)
// and the interior (main stream)? Only slab handles with a neighbour; not while per-kernel timers are on.
bool split_frame(const RtsContext* h) {
    return h->multi && h->overlap_exchange && !h->profiling && (h->SL.has[0] || h->SL.has[1]) && !h->fast && !h->SL.peer_mode && h->F.ny >= 4;
}

#define LAUNCH_DENSE(KV, MU)                                                                         \
    do {                                                                                             \
        if (h->fast) launch_chain(pdl_ok, k_step_dense_fast<KV, MU>, dgrid, dim3(128), st, h->P, h->M, h->F, Aq, SLq);   \
        else launch_chain(pdl_ok, k_step_dense<KV, MU>, dgrid, dim3(128), st, h->P, h->M, h->F, Aq, SLq);                \
    } while (0)

// arrival kernels, then — slab handles with a neighbour — the edge rows on the comm stream (their own queue of crowded
// agents is drained right behind them)
void launch_step_edges(RtsContext* h) {
    const uint32_t nb = h->multi ? h->cap_main : h->n;
    if (h->P.enable_arrival && nb && h->M.n_groups) {
        if (h->multi) k_area_sync_counts<<<blocks_for(h->M.n_groups, 256), 256, 0, h->stream>>>(h->M, 3.14159265358979323846f * h->P.rhard * h->P.rhard);
        else k_area_sync<<<blocks_for(h->M.n_groups, 256), 256, 0, h->stream>>>(h->M);
        k_arrive<<<blocks_for(nb, 256), 256, 0, h->stream>>>(nb, h->P, h->M, h->F, h->A, h->multi ? 1 : 0);
        h->launches += 2;
    }
    const bool split = nb && split_frame(h);
    h->frame_is_split = split;
    if (!split) return;
    {   // |v| <= max_speed * mul_velocity (clamp, PhysicsSystem.cpp:34-37) + seek_max_speed (SeekSystem.cpp:92)
        float vmax = 0.f;
        for (const RtsUnitType& u : h->types_host) vmax = std::max(vmax, std::fabs(u.max_speed * h->P.mul_velocity) + std::fabs(u.seek_max_speed));
        h->max_type_speed = vmax;
        set_edge_rows(h);
    }
    cudaEventRecord(h->ev_fork2, h->stream);
    cudaStreamWaitEvent(h->stream4, h->ev_fork2, 0);
    AgentArrays Aq = h->A;
    Aq.dense_q = h->A.dense_q + h->cap;
    Aq.dense_n = h->A.dense_n + 4;
    const SlabDev SLq = h->SL;
    const dim3 grid(2 * SLq.edge_blocks);
    if (h->keep_vel) k_step<true, true, false, 1><<<grid, STEP_BLOCK, 0, h->stream4>>>(nb, h->P, h->M, h->F, Aq, SLq, 0u, 0, 0u);
    else k_step<false, true, false, 1><<<grid, STEP_BLOCK, 0, h->stream4>>>(nb, h->P, h->M, h->F, Aq, SLq, 0u, 0, 0u);
    h->launches += 1;
    if (USE_DENSE_KERNEL) {
        const dim3 dgrid(h->n_sm * 2);
        cudaStream_t st = h->stream4;
        const bool pdl_ok = false;
        if (h->keep_vel) LAUNCH_DENSE(true, true);
        else LAUNCH_DENSE(false, true);
        h->launches += 1;
    }
}

// the frame kernel over everything (or, after launch_step_edges split the frame, over the interior) + k_step_dense
void launch_step_interior(RtsContext* h) {
    const uint32_t nb = h->multi ? h->cap_main : h->n;
    const uint32_t block = h->fast ? FAST_BLOCK : STEP_BLOCK;
    const dim3 grid(blocks_for(nb, block));
    const bool split = h->frame_is_split;
    const int dv = (h->keep_vel ? 2 : 0) | (h->multi ? 1 : 0);
    SlabDev SLq = h->SL;
    SLq.interior_launch = split ? 1 : 0;
    SLq.publish_here = 1;   // (read by k_step_dense* only: the last packing kernel of the frame)
    const AgentArrays& Aq = h->A;
    // leading CTAs that search the walk queue of the previous frame (see launch_scan_reorder)
    const uint32_t wc = (h->walkq_pending && nb) ? walk_ctas_of(h) : 0u;
    const int wp = h->walkq_parity;
    if (wc) h->walkq_pending = false;
    const dim3 grid_w(grid.x + wc);
    // persistent launches (every launch but the two halves of a split frame): one residency of the device, warps take
    // 32-slot tiles — the walk queue's first — from A.work (see k_step)
    const uint32_t wt = wc * (block / 32u);                   // walk tiles
    const uint32_t n_tiles = wt + blocks_for(nb, 32u);
    const uint32_t resident = (uint32_t)h->n_sm * (uint32_t)(h->fast ? FAST_MIN_BLOCKS : STEP_MIN_BLOCKS);
    const dim3 grid_p(USE_PERSISTENT_STEP ? std::max(1u, std::min(resident, blocks_for(n_tiles, block / 32u))) : grid_w.x);
    const uint32_t wq = USE_PERSISTENT_STEP ? wt : wc;        // leading walk tiles, in the launch's own tile unit
    timer_begin(h, KT_STEP);
    if (split) {   // (exact mode only, see split_frame)
        if (h->keep_vel) k_step<true, true, false, 2><<<grid_w, STEP_BLOCK, 0, h->stream>>>(nb, h->P, h->M, h->F, Aq, SLq, wc, wp, 0u);
        else k_step<false, true, false, 2><<<grid_w, STEP_BLOCK, 0, h->stream>>>(nb, h->P, h->M, h->F, Aq, SLq, wc, wp, 0u);
    } else if (nb && h->P.enable_alignment) {   // extension: single-domain handles that keep the velocity (checked by check_params)
        if (h->fast) launch_chain(use_pdl(h), k_step<true, false, true, 0, true>, grid_p, dim3(block), h->stream, nb, h->P, h->M, h->F, Aq, SLq, wq, wp, n_tiles);
        else launch_chain(use_pdl(h), k_step<true, false, false, 0, true>, grid_p, dim3(block), h->stream, nb, h->P, h->M, h->F, Aq, SLq, wq, wp, n_tiles);
    } else if (nb) {
#define LAUNCH_STEP(KV, MU, FA) launch_chain(use_pdl(h), k_step<KV, MU, FA>, grid_p, dim3(block), h->stream, nb, h->P, h->M, h->F, Aq, SLq, wq, wp, n_tiles)
        const int variant = (h->keep_vel ? 4 : 0) | (h->multi ? 2 : 0) | (h->fast ? 1 : 0);
        switch (variant) {
            case 0: LAUNCH_STEP(false, false, false); break;
            case 1: LAUNCH_STEP(false, false, true); break;
            case 2: LAUNCH_STEP(false, true, false); break;
            case 3: LAUNCH_STEP(false, true, true); break;
            case 4: LAUNCH_STEP(true, false, false); break;
            case 5: LAUNCH_STEP(true, false, true); break;
            case 6: LAUNCH_STEP(true, true, false); break;
            default: LAUNCH_STEP(true, true, true); break;
        }
#undef LAUNCH_STEP
    }
    timer_end(h, KT_STEP);
    if (nb && USE_DENSE_KERNEL) {   // the crowded agents k_step queued, one warp each
        timer_begin(h, KT_DENSE);
        const dim3 dgrid(h->n_sm * RTS_DENSE_GRID_PER_SM);
        cudaStream_t st = h->stream;
        const bool pdl_ok = use_pdl(h) && !split;
        if (dv == 0) LAUNCH_DENSE(false, false);
        else if (dv == 1) LAUNCH_DENSE(false, true);
        else if (dv == 2) LAUNCH_DENSE(true, false);
        else LAUNCH_DENSE(true, true);
        timer_end(h, KT_DENSE);
        h->launches += 1;
    }
    h->launches += 1;
}
#undef LAUNCH_DENSE

// slabs: first half of a frame — run the frame kernel (which packs migrants + halo)
int32_t frame_begin(RtsContext* h) {
    launch_step_edges(h);      // slabs: the send headers were zeroed by the previous k_scatter (or rts_slab_configure)
    launch_step_interior(h);
    return RTS_OK;
}

// the stream the exchange and the ingest of this frame go to
cudaStream_t comm_stream(const RtsContext* h) { return h->frame_is_split ? h->stream4 : h->stream; }

// slabs: what arrived is appended behind the frame kernel's slots (on the comm stream when the frame is split)
void launch_ingest(RtsContext* h) {
    const uint32_t per_dir = h->SL.cap_mig + h->SL.cap_halo;
    timer_begin(h, KT_INGEST);
    launch_chain(use_pdl(h) && !h->frame_is_split, k_ingest_remote, dim3(blocks_for(per_dir, 256), 2), dim3(256), comm_stream(h), h->cap_main, h->SL, h->F, h->A);
    timer_end(h, KT_INGEST);
    h->launches += 1;
    h->ingest_done = true;
}

// slabs: second half — ingest (unless launch_frame already did it), then scan + reorder
int32_t frame_end(RtsContext* h) {
    if (h->multi) {
        if (!h->ingest_done) launch_ingest(h);
        h->ingest_done = false;
        if (h->frame_is_split) {   // the interior (main stream) and the edge rows + exchange + ingest (comm stream) meet here
            cudaEventRecord(h->ev_comm, h->stream4);
            cudaStreamWaitEvent(h->stream, h->ev_comm, 0);
            h->frame_is_split = false;
        }
    }
    return launch_scan_reorder(h);
}

// neighbour exchange over NCCL: send[0] goes to rank-1 and lands in its recv[1]; send[1] goes to rank+1's recv[0]
int32_t exchange_nccl(RtsContext* h) {
    if (!h->comm) return h->fail(RTS_E_STATE, "exchange", "rts_comm_init has not been called");
    cudaStream_t st = comm_stream(h);
    timer_begin(h, KT_EXCHANGE);
    ncclResult_t r = h->p_ncclGroupStart();
    for (int d = 0; d < 2 && r == ncclSuccess; ++d) {
        if (!h->SL.has[d]) continue;
        const int peer = h->rank + (d == 0 ? -1 : 1);
        r = h->p_ncclSend(h->ex_send[d], h->ex_bytes, ncclUint8, peer, h->comm, st);
        if (r == ncclSuccess) r = h->p_ncclRecv(h->ex_recv[d], h->ex_bytes, ncclUint8, peer, h->comm, st);
    }
    const ncclResult_t r2 = h->p_ncclGroupEnd();
    if (r == ncclSuccess) r = r2;
    timer_end(h, KT_EXCHANGE);
    if (r != ncclSuccess) return h->fail(RTS_E_NCCL, "ncclSend/ncclRecv", h->p_ncclGetErrorString(r));
    return RTS_OK;
}

// neighbour exchange over peer memory: the packing kernels (k_step, k_step_dense*, k_halo_seed) have written their records
// into the neighbours' buffers; the first thread of k_ingest_remote, the next kernel of the stream, publishes them
// (publish_to_peers) — nothing to launch here
int32_t exchange_peer(RtsContext* h, bool) {
    h->exchange_pending = true;
    return RTS_OK;
}
// slab handles with the arrival rule: the stops each rank has claimed so far, summed over the ranks (k_area_sync_counts of the
// next frame turns the sum into the groups' finished areas). The one collective on the frame path — and only with the rule on.
int32_t reduce_stop_counts(RtsContext* h) {
    if (!h->multi || !h->P.enable_arrival || !h->M.n_groups) return RTS_OK;
    cudaStream_t st = comm_stream(h);
    if (h->nranks > 1) {
        if (!h->comm) return h->fail(RTS_E_STATE, "exchange", "rts_comm_init has not been called");
        const ncclResult_t r = h->p_ncclAllReduce(h->M.stop_local, h->M.stop_global, h->M.n_groups, ncclUint32, ncclSum, h->comm, st);
        if (r != ncclSuccess) return h->fail(RTS_E_NCCL, "ncclAllReduce", h->p_ncclGetErrorString(r));
    } else {
        CU(cudaMemcpyAsync(h->M.stop_global, h->M.stop_local, sizeof(uint32_t) * h->M.n_groups, cudaMemcpyDeviceToDevice, st));
    }
    return RTS_OK;
}
int32_t exchange(RtsContext* h, bool published_by_frame) {
    const int32_t rc = h->SL.peer_mode ? exchange_peer(h, published_by_frame) : exchange_nccl(h);
    return rc != RTS_OK ? rc : reduce_stop_counts(h);
}

int32_t launch_frame(RtsContext* h) {
    // a split slab frame enqueues edge rows -> exchange -> ingest on the (high-priority) comm stream BEFORE the interior
    // kernel is launched, so that they do not queue up behind its CTAs
    launch_step_edges(h);
    int32_t rc = RTS_OK;
    if (h->multi && h->nranks > 1 && h->frame_is_split) {
        if ((rc = exchange(h, false)) != RTS_OK) return rc;
        launch_ingest(h);
    }
    launch_step_interior(h);
    if (h->multi && h->nranks > 1 && !h->ingest_done && (rc = exchange(h, USE_DENSE_KERNEL)) != RTS_OK) return rc;
    if (h->multi && h->nranks <= 1 && (rc = reduce_stop_counts(h)) != RTS_OK) return rc;
    return frame_end(h);
}

int32_t rebin(RtsContext* h) {  // the "n" buffers hold the population in arbitrary order -> cell order
    if (h->multi) {   // every slot that k_bin does not fill is marked unused (0xFF.. = INVALID_CELL)
        const uint32_t n_slots = h->cap_main + 2 * (h->SL.cap_mig + h->SL.cap_halo);
        CU(cudaMemsetAsync(h->A.cellrank, 0xFF, sizeof(uint2) * (size_t)n_slots, h->stream));
    }
    timer_begin(h, KT_BIN);
    if (h->n) k_bin<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, h->P, h->F, h->A);
    timer_end(h, KT_BIN);
    h->launches += 1;
    return launch_scan_reorder(h);
}

// copies the non-null host arrays of `v` (count agents) into staging; nulls the staging pointer of absent fields
int32_t stage_in(RtsContext* h, const RtsAgentView* v, uint32_t count, Staging* out) {
    int32_t rc = alloc_staging(h, std::max(count, h->cap));
    if (rc != RTS_OK) return rc;
    Staging S = h->S;
#define UP(field, T, k)                                                                                        \
    if (v->field) CU(cudaMemcpyAsync(S.field, v->field, sizeof(T) * (size_t)(k) * count, cudaMemcpyHostToDevice, h->stream)); \
    else S.field = nullptr;
    UP(r, float, 2) UP(inertia, float, 2) UP(angle, float, 1) UP(angle_vel, float, 1) UP(radius, float, 1) UP(mass, float, 1)
    UP(state, uint8_t, 1) UP(player, uint8_t, 1) UP(unit_type, uint8_t, 1) UP(path_id, int32_t, 1) UP(waypoint, int32_t, 1)
    UP(r_next, float, 2) UP(gid, uint32_t, 1)
#undef UP
    S.vel = nullptr; S.tri_idx = nullptr; S.ns_cell = nullptr; S.map_cell = nullptr; S.flags = nullptr;
    *out = S;
    return RTS_OK;
}

}  // namespace

extern "C" {

void rts_default_params(RtsParams* p, int32_t map_nx, int32_t map_ny) {
    std::memset(p, 0, sizeof(*p));
    p->map_nx = map_nx; p->map_ny = map_ny; p->map_cell_size = 5.f;
    p->ns_cell_size = 60.f;
    p->ns_nx = (int32_t)((float)(map_nx * 5) / p->ns_cell_size) + 1;
    p->ns_ny = (int32_t)((float)(map_ny * 5) / p->ns_cell_size) + 1;
    p->r_max2 = 30.f;
    p->tri_nx = map_nx / 8; p->tri_ny = map_ny / 8; p->tri_cell_size = 40.f;
    p->mul_push = 0.1f; p->mul_repulse = 1.f; p->mul_scatter = 0.1f; p->mul_align = 1.f; p->mul_seek = 1.f;
    p->mul_decay = 1.f; p->mul_velocity = 2.f;
    p->dt = (float)(1. / 60.f);
    p->rhard = 3.f; p->rscatter = 50.f; p->sys_max_speed = 25.f; p->ralign = 50.f;
    p->finish_angle = 36; p->finish_radius = 50; p->finish_n_steps = 1; p->finish_cone_len = 10;
    p->enable_arrival = 0; p->walk_all = 1;
    p->slab_row_lo = 0; p->slab_row_hi = p->ns_ny;
}

size_t rts_abi_sizeof(int32_t which) {
    switch (which) {
        case 0: return sizeof(RtsParams);
        case 1: return sizeof(RtsTriangle);
        case 2: return sizeof(RtsEdge);
        case 3: return sizeof(RtsWallTable);
        case 4: return sizeof(RtsPathTable);
        case 5: return sizeof(RtsUnitType);
        case 6: return sizeof(RtsAgentView);
        case 7: return sizeof(RtsTimings);
        case 8: return sizeof(RtsAgentDelta);
        default: return 0;
    }
}

static int32_t check_params(RtsContext* h, const RtsParams* p) {
    if (!p) return h->fail(RTS_E_ARG, "params", "null");
    if (p->map_nx <= 0 || p->map_ny <= 0 || p->ns_nx <= 0 || p->ns_ny <= 0 || p->tri_nx <= 0 || p->tri_ny <= 0)
        return h->fail(RTS_E_ARG, "params", "grid dimensions must be positive");
    if (!(p->map_cell_size > 0) || !(p->ns_cell_size > 0) || !(p->tri_cell_size > 0) || !(p->r_max2 > 0))
        return h->fail(RTS_E_ARG, "params", "cell sizes and r_max2 must be positive");
    if ((uint64_t)p->ns_nx * (uint64_t)p->ns_ny >= (1ull << 29))
        return h->fail(RTS_E_ARG, "params", "physics grid has more than 2^29 cells");
    if (p->ns_nx > 8192 || p->ns_ny > 65536)   // key2.y packs the cell as 13 + 16 bits (pack_cs)
        return h->fail(RTS_E_ARG, "params", "physics grid is wider than 8192 or taller than 65536 cells");
    if (p->precision > RTS_PRECISION_FAST) return h->fail(RTS_E_ARG, "params", "unknown precision mode");
    if (p->enable_alignment && (h->multi || p->reserved[0] != 0))
        return h->fail(RTS_E_STATE, "params", "enable_alignment needs a single-domain handle that keeps the per-step velocity (halo records carry no velocity)");
    if (p->enable_alignment && !(p->ralign > 0)) return h->fail(RTS_E_ARG, "params", "ralign must be positive");
    if (p->enable_arrival && h->multi && !(p->finish_radius + 0.25f < p->ns_cell_size))
        return h->fail(RTS_E_ARG, "params", "enable_arrival on a slab handle needs finish_radius < ns_cell_size (the halo band is finish_radius wide "
                                            "and a slab keeps one physics row of its neighbour)");
    if (p->ns_cell_size * p->ns_cell_size < p->r_max2)
        return h->fail(RTS_E_ARG, "params", "ns_cell_size is smaller than the interaction range");
    return RTS_OK;
}

int32_t rts_create(const RtsParams* params, int32_t device_id, RtsHandle* out) {
    if (!out) return RTS_E_ARG;
    *out = nullptr;
    RtsContext* h = new (std::nothrow) RtsContext();
    if (!h) return RTS_E_ARG;
    int32_t rc = check_params(h, params);
    if (rc != RTS_OK) { delete h; return rc; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_id < 0 || device_id >= count) {
        delete h;
        return RTS_E_CUDA;  // no CPU fallback
    }
    h->device = device_id;
    h->P = *params;
    h->keep_vel = params->reserved[0] == 0;   // reserved[0] = 1 drops the per-step velocity output
    if (params->reserved[1]) { std::memcpy(&h->fine_cs, &params->reserved[1], sizeof(float)); h->fine_cs_auto = false; }
    h->use_graph = params->reserved[2] == 0;
    h->fast = params->precision == RTS_PRECISION_FAST;
    h->M.inv_ns_cs = 1.0f / params->ns_cell_size;
    h->M.inv_map_cs = 1.0f / params->map_cell_size;
    h->overlap_walk = params->reserved[4] == 0;
    h->overlap_exchange = (params->reserved[5] & 8u) != 0;
    h->graph_slabs = (params->reserved[5] & 16u) != 0;
    h->want_peer = (params->reserved[5] & 32u) == 0;
    h->pdl = (params->reserved[5] & 64u) == 0;
    *out = h;
    int prio_lo = 0, prio_hi = 0;   // the comm stream (edge rows, exchange, ingest of a slab) outranks the bulk kernels
    cudaSetDevice(device_id);
    if (cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device_id) != cudaSuccess || h->n_sm <= 0) h->n_sm = 148;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_scatter, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_walk, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        *out = nullptr;
        delete h;
        return RTS_E_CUDA;
    }
    rc = setup_fine_grid(h);
    if (rc != RTS_OK) return rc;
    CU(cudaMalloc(&h->n_dev, 16));
    CU(cudaMemsetAsync(h->n_dev, 0, 16, h->stream));
    // default unit types (UnitTypes.txt has two; dynamics constants are the component defaults) and an empty path table
    RtsUnitType ut[2];
    for (auto& u : ut) { u = RtsUnitType{}; u.lambda = 0.169f; u.max_speed = 25.f; u.seek_max_speed = 50.f; u.turn_rate = 5.f; }
    if ((rc = rts_upload_unit_types(h, ut, 2)) != RTS_OK) return rc;
    RtsPathTable pt{};
    uint32_t zero = 0;
    pt.n_paths = 0; pt.offsets = &zero;
    if ((rc = rts_upload_paths(h, &pt)) != RTS_OK) return rc;
    // AcosTable<1000>, src/core.h:115-124 — built on the host so both sides use the same libm values
    std::vector<float> table(1000);
    for (int bin = 0; bin < 1000; ++bin) {
        const float dot_value = 2 * static_cast<float>(bin) / 1000 - 1;
        table[bin] = 180.f / (3.14159265358979323846f) * std::acos(dot_value);
    }
    float* dt = nullptr;
    if ((rc = dev_alloc<float>(h, &dt, 1000, h->type_allocs)) != RTS_OK) return rc;
    CU(cudaMemcpyAsync(dt, table.data(), 4000, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->M.acos_table = dt;
    return RTS_OK;
}

int32_t rts_destroy(RtsHandle h) {
    if (!h) return RTS_E_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream2);
    cudaStreamSynchronize(h->stream4);
    cudaStreamSynchronize(h->stream);
    h->drop_graphs();
    free_all(h->table_allocs); free_all(h->path_allocs); free_all(h->type_allocs);
    if (h->path_arena) cudaFree(h->path_arena);
    if (h->border_dev) cudaFree(h->border_dev);
    if (h->delta_shadow) cudaFree(h->delta_shadow);
    if (h->delta_changes) cudaFree(h->delta_changes);
    if (h->delta_count) cudaFree(h->delta_count);
    if (h->delta_count_host) cudaFreeHost(h->delta_count_host);
    if (h->ev_delta) cudaEventDestroy(h->ev_delta);
    if (h->group_arena) cudaFree(h->group_arena);
    free_all(h->agent_allocs); free_all(h->staging_allocs); free_all(h->slab_allocs);
    for (auto& pa : h->peer_arena)
        if (pa) { cudaIpcCloseMemHandle(pa); pa = nullptr; }
    if (h->arena) cudaFree(h->arena);
    if (h->comm && h->p_ncclCommDestroy) h->p_ncclCommDestroy(h->comm);
    if (h->n_dev) cudaFree(h->n_dev);
    if (h->F.count) { cudaFree(h->F.count); cudaFree(h->F.start); cudaFree(h->F.partial); cudaFree(h->F.status); cudaFree(h->F.ticket); }
    if (h->scratch) cudaFree(h->scratch);
    for (auto& t : h->kt) {
        for (auto e : t.beg) cudaEventDestroy(e);
        for (auto e : t.end) cudaEventDestroy(e);
    }
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev_scatter); cudaEventDestroy(h->ev_walk);
    cudaEventDestroy(h->ev_fork2); cudaEventDestroy(h->ev_comm);
    cudaStreamDestroy(h->stream4);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return RTS_OK;
}

const char* rts_last_error(RtsHandle h) { return h ? h->err.c_str() : "null handle"; }

int32_t rts_set_params(RtsHandle h, const RtsParams* params) {
    if (!h) return RTS_E_ARG;
    int32_t rc = check_params(h, params);
    if (rc != RTS_OK) return rc;
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    const bool regrid = params->map_nx != h->P.map_nx || params->map_ny != h->P.map_ny || params->r_max2 != h->P.r_max2 ||
                        params->map_cell_size != h->P.map_cell_size || params->slab_row_lo != h->P.slab_row_lo ||
                        params->slab_row_hi != h->P.slab_row_hi || params->ns_cell_size != h->P.ns_cell_size;
    const bool rens = params->ns_nx != h->P.ns_nx || params->ns_ny != h->P.ns_ny || params->ns_cell_size != h->P.ns_cell_size;
    h->P = *params;
    h->fast = params->precision == RTS_PRECISION_FAST;
    h->M.inv_ns_cs = 1.0f / params->ns_cell_size;
    h->M.inv_map_cs = 1.0f / params->map_cell_size;
    if (h->multi) h->SL.halo = halo_band(h->P);   // (takes effect with the next frame's packing; call rts_halo_refresh after switching the arrival rule on)
    h->drop_graphs();
    if (regrid || rens) {
        CU(cudaStreamSynchronize(h->stream));
        if (regrid && (rc = setup_fine_grid(h)) != RTS_OK) return rc;
        if (h->n) {  // re-bin the population on the new grid: current state -> "n" buffers -> cell order
            Staging none{};
            k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, none, h->A);
            if ((rc = rebin(h)) != RTS_OK) return rc;
        }
    }
    return RTS_OK;
}

// Triangulation::updateCellGrid on the device (see k_cellgrid_classify): fills the handle's cell -> triangle cache from
// the uploaded triangles
int32_t build_cell_grid(RtsContext* h, uint32_t last_found, uint32_t z_lo = 0, uint32_t z_hi = 0xFFFFFFFFu) {
    const uint32_t n_cells = (uint32_t)h->P.tri_nx * (uint32_t)h->P.tri_ny;
    z_hi = std::min(z_hi, n_cells);
    if (z_lo >= z_hi) return RTS_OK;
    const uint32_t span = z_hi - z_lo;
    uint32_t* first = nullptr;
    uint8_t* amb = nullptr;
    CU(cudaMalloc(&first, sizeof(uint32_t) * (size_t)span));
    if (cudaMalloc(&amb, span) != cudaSuccess) { cudaFree(first); return h->fail(RTS_E_CUDA, "build_cell_grid", "out of device memory"); }
    const uint32_t nb = blocks_for(span, 128);
    if (h->M.coords_fit_i32) {
        k_cellgrid_classify<false><<<nb, 128, 0, h->stream>>>(h->P, h->M, h->c2t_dev, first, amb, z_lo, z_hi);
        k_cellgrid_resolve<false><<<nb, 128, 0, h->stream>>>(h->P, h->M, last_found, h->c2t_dev, first, amb, z_lo, z_hi);
    } else {
        k_cellgrid_classify<true><<<nb, 128, 0, h->stream>>>(h->P, h->M, h->c2t_dev, first, amb, z_lo, z_hi);
        k_cellgrid_resolve<true><<<nb, 128, 0, h->stream>>>(h->P, h->M, last_found, h->c2t_dev, first, amb, z_lo, z_hi);
    }
    h->launches += 2;
    const cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(first);
    cudaFree(amb);
    if (e != cudaSuccess) return h->fail(RTS_E_CUDA, "build_cell_grid", cudaGetErrorString(e));
    CU(cudaGetLastError());
    return RTS_OK;
}

extern "C++" {
namespace {
inline int4 tri_word(const RtsTriangle& T, int k) {
    if (k == 0) return make_int4(T.vx[0], T.vy[0], T.vx[1], T.vy[1]);
    if (k == 1) return make_int4(T.vx[2], T.vy[2], (int)T.neighbours[0], (int)T.neighbours[1]);
    return make_int4((int)T.neighbours[2], T.is_constrained[0] | (T.is_constrained[1] << 1) | (T.is_constrained[2] << 2) | (T.type << 8), 0, 0);
}
// cache cells a triangle's bounding box touches (clipped to the grid; x0 > x1 = entirely outside)
struct CellBox { int x0, x1, y0, y1; };
inline CellBox tri_cells(const RtsParams& P, const RtsTriangle& T) {
    auto range = [&](int32_t lo, int32_t hi, int32_t n, int& c0, int& c1) {
        c0 = std::max((int)std::floor((float)lo / P.tri_cell_size), 0);
        c1 = std::min((int)std::floor((float)hi / P.tri_cell_size), n - 1);
    };
    CellBox b;
    range(std::min({T.vx[0], T.vx[1], T.vx[2]}), std::max({T.vx[0], T.vx[1], T.vx[2]}), P.tri_nx, b.x0, b.x1);
    range(std::min({T.vy[0], T.vy[1], T.vy[2]}), std::max({T.vy[0], T.vy[1], T.vy[2]}), P.tri_ny, b.y0, b.y1);
    return b;
}
inline bool tri_is_outer(const RtsParams& P, const RtsTriangle& T) {
    const float gw = P.tri_nx * P.tri_cell_size, gh = P.tri_ny * P.tri_cell_size;
    const int32_t xmin = std::min({T.vx[0], T.vx[1], T.vx[2]}), xmax = std::max({T.vx[0], T.vx[1], T.vx[2]});
    const int32_t ymin = std::min({T.vy[0], T.vy[1], T.vy[2]}), ymax = std::max({T.vy[0], T.vy[1], T.vy[2]});
    return xmin < 0 || ymin < 0 || (float)xmax > gw || (float)ymax > gh;
}
inline bool box_is_large(const CellBox& b) { return (uint64_t)(b.x1 - b.x0 + 1) * (uint64_t)(b.y1 - b.y0 + 1) > SCAN_LARGE_CELLS; }
// wall pre-filter bits an edge sets (see may_touch_wall): every block its bounding box, grown by the reach, touches
template <class Fn>
inline void wall_blocks(const RtsParams& P, int32_t bx, int32_t by, float filter_r, float4 ft, float l, Fn fn) {
    const float wall_block = P.map_cell_size;
    const float x0 = ft.x, y0 = ft.y, x1 = x0 + ft.z * l, y1 = y0 + ft.w * l;
    const float reach = filter_r + 1.0f;
    const int ix0 = std::max(0, (int)std::floor((std::min(x0, x1) - reach) / wall_block));
    const int ix1 = std::min(bx - 1, (int)std::floor((std::max(x0, x1) + reach) / wall_block));
    const int iy0 = std::max(0, (int)std::floor((std::min(y0, y1) - reach) / wall_block));
    const int iy1 = std::min(by - 1, (int)std::floor((std::max(y0, y1) + reach) / wall_block));
    for (int iy = iy0; iy <= iy1; ++iy)
        for (int ix = ix0; ix <= ix1; ++ix) fn((size_t)iy * bx + ix);
}
}  // namespace
}  // extern "C++"

// MapDev::border_top / border_right, after every change of the triangle table or the cell -> triangle cache
int32_t build_border_cache(RtsContext* h) {
    const RtsParams& P = h->P;
    const float gwf = P.tri_nx * P.tri_cell_size, ghf = P.tri_ny * P.tri_cell_size;
    h->M.border_top = nullptr; h->M.border_right = nullptr; h->M.border_gw = 0; h->M.border_gh = 0;
    if (!(gwf == std::floor(gwf) && ghf == std::floor(ghf) && gwf >= 1.f && ghf >= 1.f && gwf < 1.0e6f && ghf < 1.0e6f) || !h->M.n_tris) return RTS_OK;
    const int32_t gw = (int32_t)gwf, gh = (int32_t)ghf;
    const size_t need = (size_t)gw + 1 + (size_t)gh + 1;
    if (need > h->border_cap) {
        if (h->border_dev) cudaFree(h->border_dev);
        h->border_dev = nullptr; h->border_cap = 0;
        CU(cudaMalloc(&h->border_dev, need * sizeof(uint32_t)));
        h->border_cap = need;
    }
    uint32_t* top = h->border_dev;
    uint32_t* right = h->border_dev + gw + 1;
    const dim3 grid(blocks_for((uint32_t)need, 128));
    if (h->M.coords_fit_i32) k_border_cache<false><<<grid, 128, 0, h->stream>>>(P, h->M, top, right, gw, gh);
    else k_border_cache<true><<<grid, 128, 0, h->stream>>>(P, h->M, top, right, gw, gh);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    h->M.border_top = top; h->M.border_right = right; h->M.border_gw = gw; h->M.border_gh = gh;
    return RTS_OK;
}

int32_t rts_upload_map(RtsHandle h, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri, const RtsWallTable* walls) {
    if (!h) return RTS_E_ARG;
    if (!tris || !n_tris || !walls) return h->fail(RTS_E_ARG, "rts_upload_map", "null table");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    CU(cudaStreamSynchronize(h->stream));
    free_all(h->table_allocs);
    h->map_ready = false;
    const RtsParams& P = h->P;
    uint64_t bytes = 0;
    const uint32_t expect[4] = {(uint32_t)P.map_ny + 1, (uint32_t)(P.map_nx + P.map_ny) + 1, (uint32_t)P.map_nx + 1,
                                (uint32_t)(P.map_nx + P.map_ny) + 1};
    for (int o = 0; o < 4; ++o) {
        if (!walls->line_offsets[o]) return h->fail(RTS_E_ARG, "rts_upload_map", "null line_offsets");
        if (walls->n_lines[o] != expect[o]) return h->fail(RTS_E_ARG, "rts_upload_map", "n_lines does not match the MapGrid (Edges.cpp:6-19)");
    }
    // triangles -> 3 x int4 (device capacity with slack: rts_update_map_region appends the triangles a CDT insert creates)
    std::vector<int4> t4(3 * (size_t)n_tris);
    for (uint32_t t = 0; t < n_tris; ++t) {
        const RtsTriangle& T = tris[t];
        for (int k = 0; k < 3; ++k)
            if (T.neighbours[k] != 0xFFFFFFFFu && T.neighbours[k] >= n_tris) return h->fail(RTS_E_ARG, "rts_upload_map", "triangle neighbour out of range");
        for (int k = 0; k < 3; ++k) t4[3 * t + k] = tri_word(T, k);
    }
    const size_t n_tri_cells = (size_t)P.tri_nx * P.tri_ny;
    for (size_t c = 0; cell2tri && c < n_tri_cells; ++c)
        if (cell2tri[c] != 0xFFFFFFFFu && cell2tri[c] >= n_tris) return h->fail(RTS_E_ARG, "rts_upload_map", "cell2tri entry out of range");
    h->tris_host.assign(tris, tris + n_tris);
    h->tri_cap = n_tris + n_tris / 2 + 4096;
    int4* d_tris = nullptr; uint32_t* d_c2t = nullptr;
    int32_t rc;
    if ((rc = dev_alloc(h, &d_tris, 3 * (size_t)h->tri_cap, h->table_allocs)) != RTS_OK) return rc;
    if ((rc = dev_alloc(h, &d_c2t, n_tri_cells, h->table_allocs)) != RTS_OK) return rc;
    CU(cudaMemcpyAsync(d_tris, t4.data(), t4.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    bytes += t4.size() * sizeof(int4);
    if (cell2tri) { CU(cudaMemcpyAsync(d_c2t, cell2tri, n_tri_cells * 4, cudaMemcpyHostToDevice, h->stream)); bytes += n_tri_cells * 4; }
    // walls -> per line {first, end} into one edge pool
    std::vector<uint2> se;
    std::vector<float4> ft;
    std::vector<float> el;
    MapDev& M = h->M;
    for (int o = 0; o < 4; ++o) {
        M.line_base[o] = (uint32_t)se.size();
        M.n_lines[o] = walls->n_lines[o];
        const uint32_t* lo = walls->line_offsets[o];
        const uint32_t base = (uint32_t)ft.size();
        for (uint32_t l = 0; l < walls->n_lines[o]; ++l) {
            if (lo[l + 1] < lo[l]) return h->fail(RTS_E_ARG, "rts_upload_map", "line_offsets not monotone");
            se.push_back(make_uint2(base + lo[l], base + lo[l + 1]));
        }
        const uint32_t ne = lo[walls->n_lines[o]];
        if (ne && !walls->edges[o]) return h->fail(RTS_E_ARG, "rts_upload_map", "null edges");
        for (uint32_t e = 0; e < ne; ++e) {
            const RtsEdge& E = walls->edges[o][e];
            ft.push_back(make_float4(E.from_x, E.from_y, E.t_x, E.t_y));
            el.push_back(E.l);
        }
    }
    h->line_se_host = se;
    h->edge_pool_used = (uint32_t)ft.size();
    h->edge_pool_cap = h->edge_pool_used + h->edge_pool_used / 2 + 4096;
    uint2* d_se = nullptr; float4* d_ft = nullptr; float* d_el = nullptr;
    if ((rc = dev_alloc(h, &d_se, se.size(), h->table_allocs)) != RTS_OK) return rc;
    if ((rc = dev_alloc(h, &d_ft, h->edge_pool_cap, h->table_allocs)) != RTS_OK) return rc;
    if ((rc = dev_alloc(h, &d_el, h->edge_pool_cap, h->table_allocs)) != RTS_OK) return rc;
    CU(cudaMemcpyAsync(d_se, se.data(), se.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
    bytes += se.size() * sizeof(uint2);
    if (!ft.empty()) {
        CU(cudaMemcpyAsync(d_ft, ft.data(), ft.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(d_el, el.data(), el.size() * 4, cudaMemcpyHostToDevice, h->stream));
        bytes += ft.size() * 20;
    }
    // conservative wall pre-filter bitmap (see may_touch_wall): one bit per MapGrid cell, set if a wall can lie within
    // filter_r (+1) of any point of the cell — the bounding box of every wall, grown by that reach
    const float filter_r = 8.0f;
    const int32_t bx = (int32_t)std::ceil(P.map_nx * P.map_cell_size / P.map_cell_size) + 1;
    const int32_t by = (int32_t)std::ceil(P.map_ny * P.map_cell_size / P.map_cell_size) + 1;
    std::vector<uint32_t>& bits = h->wall_bits_host;
    bits.assign(((size_t)bx * by + 31) / 32, 0u);
    for (size_t e = 0; e < ft.size(); ++e) wall_blocks(P, bx, by, filter_r, ft[e], el[e], [&](size_t b) { bits[b >> 5] |= 1u << (b & 31); });
    uint32_t* d_bits = nullptr;
    if ((rc = dev_alloc(h, &d_bits, bits.size(), h->table_allocs)) != RTS_OK) return rc;
    CU(cudaMemcpyAsync(d_bits, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, h->stream));
    bytes += bits.size() * 4;
    // scan lists for findTriangle's linear-scan fallback (see MapDev): per cache cell the ascending list of the small triangles
    // whose bounding box touches it
    std::vector<uint32_t> scan_off(n_tri_cells + 1, 0u), scan_list, scan_large, scan_outer;
    {
        std::vector<CellBox> box(n_tris);
        std::vector<uint8_t> large(n_tris, 0);
        for (uint32_t t = 0; t < n_tris; ++t) {
            const RtsTriangle& T = tris[t];
            if (tri_is_outer(P, T)) scan_outer.push_back(t);
            box[t] = tri_cells(P, T);
            if (box[t].x0 > box[t].x1 || box[t].y0 > box[t].y1) continue;   // entirely outside the grid
            if (box_is_large(box[t])) { large[t] = 1; scan_large.push_back(t); continue; }
            for (int y = box[t].y0; y <= box[t].y1; ++y)
                for (int x = box[t].x0; x <= box[t].x1; ++x) scan_off[(size_t)y * P.tri_nx + x + 1]++;
        }
        for (size_t c = 0; c < n_tri_cells; ++c) scan_off[c + 1] += scan_off[c];
        scan_list.resize(scan_off[n_tri_cells]);
        std::vector<uint32_t> fill(scan_off.begin(), scan_off.end() - 1);
        for (uint32_t t = 0; t < n_tris; ++t) {   // ascending t => every cell list is ascending
            if (large[t] || box[t].x0 > box[t].x1 || box[t].y0 > box[t].y1) continue;
            for (int y = box[t].y0; y <= box[t].y1; ++y)
                for (int x = box[t].x0; x <= box[t].x1; ++x) scan_list[fill[(size_t)y * P.tri_nx + x]++] = t;
        }
    }
    h->scan_se_host.resize(n_tri_cells);
    for (size_t c = 0; c < n_tri_cells; ++c) h->scan_se_host[c] = make_uint2(scan_off[c], scan_off[c + 1]);
    h->scan_pool_used = (uint32_t)scan_list.size();
    h->scan_pool_cap = h->scan_pool_used + h->scan_pool_used / 2 + 65536;
    h->scan_large_cap = (uint32_t)scan_large.size() + 1024;
    h->scan_outer_cap = (uint32_t)scan_outer.size() + 1024;
    uint2* d_sse = nullptr; uint32_t *d_slist = nullptr, *d_slarge = nullptr, *d_souter = nullptr;
    if ((rc = dev_alloc(h, &d_souter, h->scan_outer_cap, h->table_allocs)) != RTS_OK) return rc;
    if (!scan_outer.empty()) CU(cudaMemcpyAsync(d_souter, scan_outer.data(), scan_outer.size() * 4, cudaMemcpyHostToDevice, h->stream));
    if ((rc = dev_alloc(h, &d_sse, n_tri_cells, h->table_allocs)) != RTS_OK) return rc;
    if ((rc = dev_alloc(h, &d_slist, h->scan_pool_cap, h->table_allocs)) != RTS_OK) return rc;
    if ((rc = dev_alloc(h, &d_slarge, h->scan_large_cap, h->table_allocs)) != RTS_OK) return rc;
    CU(cudaMemcpyAsync(d_sse, h->scan_se_host.data(), n_tri_cells * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
    if (!scan_list.empty()) CU(cudaMemcpyAsync(d_slist, scan_list.data(), scan_list.size() * 4, cudaMemcpyHostToDevice, h->stream));
    if (!scan_large.empty()) CU(cudaMemcpyAsync(d_slarge, scan_large.data(), scan_large.size() * 4, cudaMemcpyHostToDevice, h->stream));
    bytes += n_tri_cells * sizeof(uint2) + (scan_list.size() + scan_large.size() + scan_outer.size()) * 4;
    // int32 triangle walk is exact iff no product can overflow: |coord| <= 16383 (see tri_sign)
    int32_t cmax = 0;
    for (uint32_t t = 0; t < n_tris; ++t)
        for (int k = 0; k < 3; ++k) cmax = std::max(cmax, std::max(std::abs(tris[t].vx[k]), std::abs(tris[t].vy[k])));
    CU(cudaStreamSynchronize(h->stream));
    M.tris = d_tris; M.n_tris = n_tris; M.cell2tri = d_c2t;
    M.line_se = d_se; M.edge_ft = d_ft; M.edge_l = d_el;
    M.wall_bits = d_bits; M.wall_bx = bx; M.wall_by = by; M.wall_filter_radius = filter_r; M.inv_wall_block = 1.0f / P.map_cell_size;
    M.coords_fit_i32 = cmax <= 16383 ? 1 : 0;
    h->A.n_tris_dev_hint = n_tris;
    M.scan_se = d_sse; M.scan_list = d_slist; M.scan_large = d_slarge; M.n_scan_large = (uint32_t)scan_large.size();
    M.scan_outer = d_souter; M.n_scan_outer = (uint32_t)scan_outer.size();
    h->c2t_dev = d_c2t;
    h->tris_w = d_tris; h->line_se_w = d_se; h->edge_ft_w = d_ft; h->edge_l_w = d_el; h->scan_se_w = d_sse; h->scan_list_w = d_slist;
    h->scan_large_w = d_slarge; h->scan_outer_w = d_souter; h->wall_bits_w = d_bits;
    h->upload_bytes_last = h->upload_bytes_full = bytes;
    if (!cell2tri && (rc = build_cell_grid(h, 0u)) != RTS_OK) return rc;   // no host cache given: built here
    if ((rc = build_border_cache(h)) != RTS_OK) return rc;
    h->map_ready = true;
    h->drop_graphs();
    if (h->n) {   // agents uploaded before the map: give them their triangle index now
        Staging none{};
        k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, none, h->A);
        if ((rc = rebin(h)) != RTS_OK) return rc;
        CU(cudaStreamSynchronize(h->stream));
    }
    return RTS_OK;
}

// Game::updateTriangulation after ONE building insert (Game.cpp:38-48, MapGrid.cpp:1095-1312): the CDT changed in a small
// neighbourhood (a few dozen triangle records rewritten or appended), eight walls joined eight lines. Only that travels.
int32_t rts_update_map_region(RtsHandle h, uint32_t n_tris_total, const uint32_t* tri_indices, const RtsTriangle* tri_records, uint32_t n_changed,
                              const uint32_t* line_orient, const uint32_t* line_index, const uint32_t* line_counts, const RtsEdge* line_edges,
                              uint32_t n_lines_changed, uint32_t last_found_tri) {
    if (!h) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_update_map_region", "rts_upload_map has not been called");
    if ((n_changed && (!tri_indices || !tri_records)) || (n_lines_changed && (!line_orient || !line_index || !line_counts)))
        return h->fail(RTS_E_ARG, "rts_update_map_region", "null array");
    const RtsParams& P = h->P;
    MapDev& M = h->M;
    const uint32_t n_old = M.n_tris;
    if (n_tris_total < n_old) return h->fail(RTS_E_ARG, "rts_update_map_region", "the triangle table cannot shrink (send the whole map)");
    if (n_tris_total > h->tri_cap) return h->fail(RTS_E_CAPACITY, "rts_update_map_region", "triangle capacity exhausted: call rts_upload_map");
    // ---- validate, then patch the host mirror of the triangles; the cache cells whose lists / entries can change are those
    //      the bounding boxes of the old and the new records touch
    std::vector<uint8_t> seen_new(n_tris_total - n_old, 0);
    for (uint32_t k = 0; k < n_changed; ++k) {
        if (tri_indices[k] >= n_tris_total) return h->fail(RTS_E_ARG, "rts_update_map_region", "triangle index out of range");
        if (tri_indices[k] >= n_old) seen_new[tri_indices[k] - n_old] = 1;
        for (int e = 0; e < 3; ++e)
            if (tri_records[k].neighbours[e] != 0xFFFFFFFFu && tri_records[k].neighbours[e] >= n_tris_total)
                return h->fail(RTS_E_ARG, "rts_update_map_region", "triangle neighbour out of range");
    }
    for (uint8_t v : seen_new)
        if (!v) return h->fail(RTS_E_ARG, "rts_update_map_region", "every appended triangle must be among the changed records");
    uint64_t edges_new = 0;
    for (uint32_t k = 0; k < n_lines_changed; ++k) {
        if (line_orient[k] > 3 || line_index[k] >= M.n_lines[line_orient[k]]) return h->fail(RTS_E_ARG, "rts_update_map_region", "wall line out of range");
        edges_new += line_counts[k];
    }
    if (edges_new && !line_edges) return h->fail(RTS_E_ARG, "rts_update_map_region", "null edges");
    if (h->edge_pool_used + edges_new > h->edge_pool_cap) return h->fail(RTS_E_CAPACITY, "rts_update_map_region", "wall pool exhausted: call rts_upload_map");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    CU(cudaStreamSynchronize(h->stream));
    uint64_t bytes = 0;
    int rx0 = P.tri_nx, rx1 = -1, ry0 = P.tri_ny, ry1 = -1;   // affected cache cells
    bool any_large = false, any_outer = false;
    auto grow = [&](const RtsTriangle& T) {
        const CellBox b = tri_cells(P, T);
        if (tri_is_outer(P, T)) any_outer = true;
        if (b.x0 > b.x1 || b.y0 > b.y1) return;
        if (box_is_large(b)) { any_large = true; return; }
        rx0 = std::min(rx0, b.x0); rx1 = std::max(rx1, b.x1); ry0 = std::min(ry0, b.y0); ry1 = std::max(ry1, b.y1);
    };
    h->tris_host.resize(n_tris_total);
    std::vector<int4> t4(3 * (size_t)n_changed);
    for (uint32_t k = 0; k < n_changed; ++k) {
        const uint32_t t = tri_indices[k];
        if (t < n_old) grow(h->tris_host[t]);
        h->tris_host[t] = tri_records[k];
        grow(tri_records[k]);
        for (int w = 0; w < 3; ++w) t4[3 * (size_t)k + w] = tri_word(tri_records[k], w);
    }
    // staging in the generic scratch buffer
    auto stage = [&](const void* src, size_t nbytes, void** dev) -> int32_t {
        *dev = nullptr;
        if (!nbytes) return RTS_OK;
        void* d = nullptr;
        if (cudaMalloc(&d, nbytes) != cudaSuccess) return h->fail(RTS_E_CUDA, "rts_update_map_region", "out of device memory");
        if (cudaMemcpyAsync(d, src, nbytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { cudaFree(d); return h->fail(RTS_E_CUDA, "rts_update_map_region", "copy failed"); }
        bytes += nbytes;
        *dev = d;
        return RTS_OK;
    };
    std::vector<void*> temps;
    auto staged = [&](const void* src, size_t nbytes) -> void* {
        void* d = nullptr;
        if (stage(src, nbytes, &d) == RTS_OK && d) temps.push_back(d);
        return d;
    };
    int32_t rc = RTS_OK;
    // ---- triangles
    if (n_changed) {
        uint32_t* d_idx = static_cast<uint32_t*>(staged(tri_indices, 4 * (size_t)n_changed));
        int4* d_val = static_cast<int4*>(staged(t4.data(), t4.size() * sizeof(int4)));
        if (!d_idx || !d_val) rc = RTS_E_CUDA;
        else k_scatter_tris<<<blocks_for(n_changed, 128), 128, 0, h->stream>>>(n_changed, d_idx, d_val, h->tris_w);
    }
    // ---- walls: every changed line is appended at the end of the pool and the line redirected; new wall bits are OR-ed in
    if (rc == RTS_OK && n_lines_changed) {
        std::vector<uint32_t> se_idx;
        std::vector<uint2> se_val;
        std::vector<float4> ft;
        std::vector<float> el;
        std::vector<uint32_t> bit_idx, bit_val;
        const RtsEdge* E = line_edges;
        const uint32_t first = h->edge_pool_used;
        for (uint32_t k = 0; k < n_lines_changed; ++k) {
            const uint32_t line = M.line_base[line_orient[k]] + line_index[k];
            const uint32_t beg = first + (uint32_t)ft.size();
            for (uint32_t e = 0; e < line_counts[k]; ++e, ++E) {
                const float4 f = make_float4(E->from_x, E->from_y, E->t_x, E->t_y);
                ft.push_back(f);
                el.push_back(E->l);
                wall_blocks(P, M.wall_bx, M.wall_by, M.wall_filter_radius, f, E->l, [&](size_t b) {
                    const uint32_t bit = 1u << (b & 31);
                    if (!(h->wall_bits_host[b >> 5] & bit)) { h->wall_bits_host[b >> 5] |= bit; bit_idx.push_back((uint32_t)(b >> 5)); bit_val.push_back(bit); }
                });
            }
            se_idx.push_back(line);
            se_val.push_back(make_uint2(beg, beg + line_counts[k]));
            h->line_se_host[line] = se_val.back();
        }
        if (!ft.empty()) {
            CU(cudaMemcpyAsync(h->edge_ft_w + first, ft.data(), ft.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
            CU(cudaMemcpyAsync(h->edge_l_w + first, el.data(), el.size() * 4, cudaMemcpyHostToDevice, h->stream));
            bytes += ft.size() * 20;
        }
        h->edge_pool_used += (uint32_t)ft.size();
        uint32_t* d_idx = static_cast<uint32_t*>(staged(se_idx.data(), 4 * se_idx.size()));
        uint2* d_val = static_cast<uint2*>(staged(se_val.data(), 8 * se_val.size()));
        if (!d_idx || !d_val) rc = RTS_E_CUDA;
        else k_scatter_u2<<<blocks_for((uint32_t)se_idx.size(), 128), 128, 0, h->stream>>>((uint32_t)se_idx.size(), d_idx, d_val, h->line_se_w);
        if (rc == RTS_OK && !bit_idx.empty()) {
            {
                // several bits of one word may be new: merged on the host, one OR per word on the device
                std::vector<uint32_t> widx, wval;
                for (size_t i = 0; i < bit_idx.size(); ++i) {
                    size_t j = 0;
                    for (; j < widx.size(); ++j) if (widx[j] == bit_idx[i]) break;
                    if (j == widx.size()) { widx.push_back(bit_idx[i]); wval.push_back(0u); }
                    wval[j] |= bit_val[i];
                }
                uint32_t* d_wi = static_cast<uint32_t*>(staged(widx.data(), 4 * widx.size()));
                uint32_t* d_wv = static_cast<uint32_t*>(staged(wval.data(), 4 * wval.size()));
                if (!d_wi || !d_wv) rc = RTS_E_CUDA;
                else k_scatter_u32<<<blocks_for((uint32_t)widx.size(), 128), 128, 0, h->stream>>>((uint32_t)widx.size(), d_wi, d_wv, h->wall_bits_w, 1);
            }
        }
    }
    // ---- scan lists of the affected cells, rebuilt from the mirror and appended to the pool
    if (rc == RTS_OK && rx0 <= rx1 && ry0 <= ry1) {
        const int w = rx1 - rx0 + 1, hgt = ry1 - ry0 + 1;
        std::vector<std::vector<uint32_t>> lists((size_t)w * hgt);
        for (uint32_t t = 0; t < n_tris_total; ++t) {   // ascending t => ascending lists
            const CellBox b = tri_cells(P, h->tris_host[t]);
            if (b.x0 > b.x1 || b.y0 > b.y1 || box_is_large(b)) continue;
            if (b.x1 < rx0 || b.x0 > rx1 || b.y1 < ry0 || b.y0 > ry1) continue;
            for (int y = std::max(b.y0, ry0); y <= std::min(b.y1, ry1); ++y)
                for (int x = std::max(b.x0, rx0); x <= std::min(b.x1, rx1); ++x) lists[(size_t)(y - ry0) * w + (x - rx0)].push_back(t);
        }
        std::vector<uint32_t> flat, se_idx;
        std::vector<uint2> se_val;
        for (int y = 0; y < hgt; ++y)
            for (int x = 0; x < w; ++x) {
                const auto& L = lists[(size_t)y * w + x];
                const uint32_t cell = (uint32_t)(ry0 + y) * (uint32_t)P.tri_nx + (uint32_t)(rx0 + x);
                const uint32_t beg = h->scan_pool_used + (uint32_t)flat.size();
                flat.insert(flat.end(), L.begin(), L.end());
                se_idx.push_back(cell);
                se_val.push_back(make_uint2(beg, beg + (uint32_t)L.size()));
            }
        if (h->scan_pool_used + flat.size() > h->scan_pool_cap) rc = h->fail(RTS_E_CAPACITY, "rts_update_map_region", "scan-list pool exhausted: call rts_upload_map");
        else {
            if (!flat.empty()) {
                CU(cudaMemcpyAsync(h->scan_list_w + h->scan_pool_used, flat.data(), 4 * flat.size(), cudaMemcpyHostToDevice, h->stream));
                bytes += 4 * flat.size();
            }
            h->scan_pool_used += (uint32_t)flat.size();
            for (size_t i = 0; i < se_idx.size(); ++i) h->scan_se_host[se_idx[i]] = se_val[i];
            uint32_t* d_idx = static_cast<uint32_t*>(staged(se_idx.data(), 4 * se_idx.size()));
            uint2* d_val = static_cast<uint2*>(staged(se_val.data(), 8 * se_val.size()));
            if (!d_idx || !d_val) rc = RTS_E_CUDA;
            else k_scatter_u2<<<blocks_for((uint32_t)se_idx.size(), 128), 128, 0, h->stream>>>((uint32_t)se_idx.size(), d_idx, d_val, h->scan_se_w);
        }
    }
    // ---- the few large / outer triangles: their lists are rebuilt whole when one of them changed
    if (rc == RTS_OK && (any_large || any_outer)) {
        std::vector<uint32_t> large, outer;
        for (uint32_t t = 0; t < n_tris_total; ++t) {
            const RtsTriangle& T = h->tris_host[t];
            if (tri_is_outer(P, T)) outer.push_back(t);
            const CellBox b = tri_cells(P, T);
            if (!(b.x0 > b.x1 || b.y0 > b.y1) && box_is_large(b)) large.push_back(t);
        }
        if (large.size() > h->scan_large_cap || outer.size() > h->scan_outer_cap) rc = h->fail(RTS_E_CAPACITY, "rts_update_map_region", "large-triangle list exhausted: call rts_upload_map");
        else {
            if (!large.empty()) CU(cudaMemcpyAsync(h->scan_large_w, large.data(), 4 * large.size(), cudaMemcpyHostToDevice, h->stream));
            if (!outer.empty()) CU(cudaMemcpyAsync(h->scan_outer_w, outer.data(), 4 * outer.size(), cudaMemcpyHostToDevice, h->stream));
            bytes += 4 * (large.size() + outer.size());
            M.n_scan_large = (uint32_t)large.size();
            M.n_scan_outer = (uint32_t)outer.size();
        }
    }
    if (rc == RTS_OK) {
        int32_t cmax = M.coords_fit_i32 ? 0 : 16384;
        for (uint32_t k = 0; k < n_changed; ++k)
            for (int e = 0; e < 3; ++e) cmax = std::max(cmax, std::max(std::abs(tri_records[k].vx[e]), std::abs(tri_records[k].vy[e])));
        M.coords_fit_i32 = cmax <= 16383 ? 1 : 0;
        M.n_tris = n_tris_total;
        h->A.n_tris_dev_hint = n_tris_total;
    }
    const cudaError_t ce = cudaStreamSynchronize(h->stream);
    for (void* t : temps) cudaFree(t);
    if (rc != RTS_OK) return rc;
    if (ce != cudaSuccess) return h->fail(RTS_E_CUDA, "rts_update_map_region", cudaGetErrorString(ce));
    // ---- Triangulation::updateCellGrid for the band of cache rows around the change (one more row on either side: a run of
    //      ambiguous cells is resolved from its predecessor in zig-zag order)
    if (rx0 <= rx1 && ry0 <= ry1) {
        const uint32_t j0 = (uint32_t)std::max(ry0 - 1, 0), j1 = (uint32_t)std::min(ry1 + 1, P.tri_ny - 1);
        if ((rc = build_cell_grid(h, last_found_tri, j0 * (uint32_t)P.tri_nx, (j1 + 1) * (uint32_t)P.tri_nx)) != RTS_OK) return rc;
    }
    if (any_large && (rc = build_cell_grid(h, last_found_tri)) != RTS_OK) return rc;   // a triangle spanning > 1024 cells changed: whole grid
    if ((rc = build_border_cache(h)) != RTS_OK) return rc;   // (the walk across the map may pass through the changed triangles)
    h->upload_bytes_last = bytes;
    h->drop_graphs();
    return RTS_OK;
}

int32_t rts_get_cell_grid(RtsHandle h, uint32_t* cell2tri_out) {
    if (!h || !cell2tri_out) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_get_cell_grid", "rts_upload_map has not been called");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(cell2tri_out, h->c2t_dev, sizeof(uint32_t) * (size_t)h->P.tri_nx * h->P.tri_ny, cudaMemcpyDeviceToHost));
    return RTS_OK;
}

int32_t rts_map_upload_bytes(RtsHandle h, uint64_t* last_call, uint64_t* last_full_upload) {
    if (!h) return RTS_E_ARG;
    if (last_call) *last_call = h->upload_bytes_last;
    if (last_full_upload) *last_full_upload = h->upload_bytes_full;
    return RTS_OK;
}

int32_t rts_upload_unit_types(RtsHandle h, const RtsUnitType* types, uint32_t n_types) {
    if (!h) return RTS_E_ARG;
    if (!types || n_types == 0 || n_types > 256) return h->fail(RTS_E_ARG, "rts_upload_unit_types", "need 1..256 types");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    CU(cudaStreamSynchronize(h->stream));
    RtsUnitType* d = nullptr;
    void* old = const_cast<RtsUnitType*>(h->M.unit_types);
    int32_t rc = dev_alloc(h, &d, 256, h->type_allocs);
    if (rc != RTS_OK) return rc;
    std::vector<RtsUnitType> full(256, types[0]);   // unknown type ids behave like type 0
    std::memcpy(full.data(), types, sizeof(RtsUnitType) * n_types);
    CU(cudaMemcpyAsync(d, full.data(), sizeof(RtsUnitType) * 256, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->M.unit_types = d; h->M.n_types = n_types;
    h->types_host.assign(types, types + n_types);
    if (old) {
        for (auto it = h->type_allocs.begin(); it != h->type_allocs.end(); ++it)
            if (*it == old) { cudaFree(old); h->type_allocs.erase(it); break; }
    }
    h->drop_graphs();
    return RTS_OK;
}

int32_t rts_upload_paths(RtsHandle h, const RtsPathTable* paths) {
    if (!h) return RTS_E_ARG;
    if (!paths || !paths->offsets) return h->fail(RTS_E_ARG, "rts_upload_paths", "null table");
    const uint32_t np = paths->n_paths;
    const uint32_t nw = paths->offsets[np];
    if (np && (!paths->waypoints || !paths->portals || !paths->path_end)) return h->fail(RTS_E_ARG, "rts_upload_paths", "null arrays");
    for (uint32_t p = 0; p < np; ++p)
        if (paths->offsets[p + 1] <= paths->offsets[p]) return h->fail(RTS_E_ARG, "rts_upload_paths", "every path needs at least one waypoint");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    CU(cudaStreamSynchronize(h->stream));
    const uint32_t ng = paths->group ? paths->n_groups : np;
    std::vector<int32_t> grp(std::max<uint32_t>(np, 1), 0);
    for (uint32_t p = 0; p < np; ++p) {
        grp[p] = paths->group ? paths->group[p] : (int32_t)p;
        if (grp[p] < 0 || (uint32_t)grp[p] >= ng) return h->fail(RTS_E_ARG, "rts_upload_paths", "group id out of range");
    }
    // One device arena for the tables (grown by half when it gets too small, otherwise reused: the adapter re-sends its
    // per-agent windows whenever the planner has rewritten some, and eleven cudaMalloc / cudaFree per call cost more than the
    // frame), filled from one host image with one copy. The groups' arrival state (finished areas, stop counts) lives in an
    // arena of its own that survives every re-upload with the same number of groups.
    auto align256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_hdr = 0;
    const size_t o_grp = align256(o_hdr + sizeof(float4) * std::max<uint32_t>(np, 1));
    const size_t o_off = align256(o_grp + sizeof(int32_t) * std::max<uint32_t>(np, 1));
    const size_t o_wp = align256(o_off + sizeof(uint32_t) * ((size_t)np + 1));
    const size_t o_end = align256(o_wp + sizeof(float4) * 2 * std::max<size_t>(nw, 1));
    const size_t need = align256(o_end + sizeof(float2) * std::max<uint32_t>(np, 1));
    if (need > h->path_arena_cap) {
        if (h->path_arena) cudaFree(h->path_arena);
        h->path_arena = nullptr; h->path_arena_cap = 0;
        const size_t cap = need + need / 2;
        CU(cudaMalloc(&h->path_arena, cap));
        h->path_arena_cap = cap;
    }
    std::vector<char>& img = h->path_image;
    img.assign(need, 0);
    {
        float4* hdr = reinterpret_cast<float4*>(img.data() + o_hdr);   // {offset, end offset (as bits), path_end}: one load per agent in the frame kernel
        for (uint32_t p = 0; p < np; ++p) {
            float4 w;
            std::memcpy(&w.x, &paths->offsets[p], 4);
            std::memcpy(&w.y, &paths->offsets[p + 1], 4);
            w.z = paths->path_end[2 * p];
            w.w = paths->path_end[2 * p + 1];
            hdr[p] = w;
        }
        std::memcpy(img.data() + o_grp, grp.data(), sizeof(int32_t) * np);
        std::memcpy(img.data() + o_off, paths->offsets, sizeof(uint32_t) * ((size_t)np + 1));
        float4* wp = reinterpret_cast<float4*>(img.data() + o_wp);
        for (uint32_t k = 0; k < nw; ++k) {
            const RtsEdge& e = paths->portals[k];
            wp[2 * k] = make_float4(paths->waypoints[2 * k], paths->waypoints[2 * k + 1], e.from_x, e.from_y);
            wp[2 * k + 1] = make_float4(e.t_x, e.t_y, e.l, 0.f);
        }
        if (np) std::memcpy(img.data() + o_end, paths->path_end, sizeof(float2) * np);
    }
    CU(cudaMemcpyAsync(h->path_arena, img.data(), need, cudaMemcpyHostToDevice, h->stream));
    if (!h->group_arena || h->group_arena_ng != ng) {   // finished areas and stop counts start from zero for a new set of groups
        if (h->group_arena) cudaFree(h->group_arena);
        h->group_arena = nullptr;
        const size_t per = align256(sizeof(float) * std::max<uint32_t>(ng, 1));
        CU(cudaMalloc(&h->group_arena, 5 * per));
        CU(cudaMemsetAsync(h->group_arena, 0, 5 * per, h->stream));
        h->group_arena_ng = ng;
    }
    CU(cudaStreamSynchronize(h->stream));   // (the host image may be rewritten by the next call)
    const MapDev before = h->M;
    {
        char* base = static_cast<char*>(h->path_arena);
        h->M.path_hdr = reinterpret_cast<float4*>(base + o_hdr);
        h->M.path_group = reinterpret_cast<int32_t*>(base + o_grp);
        h->M.path_off = reinterpret_cast<uint32_t*>(base + o_off);
        h->M.waypoints = reinterpret_cast<float4*>(base + o_wp);
        h->M.path_end = reinterpret_cast<float2*>(base + o_end);
        h->M.n_paths = np;
        h->M.n_groups = ng;
        char* gb = static_cast<char*>(h->group_arena);
        const size_t per = align256(sizeof(float) * std::max<uint32_t>(ng, 1));
        h->M.area_read = reinterpret_cast<float*>(gb);
        h->M.area_acc = reinterpret_cast<float*>(gb + per);
        h->M.stop_local = reinterpret_cast<uint32_t*>(gb + 2 * per);
        h->M.stop_global = reinterpret_cast<uint32_t*>(gb + 3 * per);
        h->M.stop_applied = reinterpret_cast<uint32_t*>(gb + 4 * per);
    }
    // (a CUDA graph holds the table pointers and sizes by value: it stays valid when only the contents changed)
    if (std::memcmp(&before, &h->M, sizeof(MapDev)) == 0) return RTS_OK;
    h->drop_graphs();
    return RTS_OK;
}

static int32_t ingest(RtsContext* h, uint32_t first, uint32_t m, const RtsAgentView* agents) {
    Staging S;
    int32_t rc = stage_in(h, agents, m, &S);
    if (rc != RTS_OK) return rc;
    if (m) k_ingest<<<blocks_for(m, 256), 256, 0, h->stream>>>(m, first, S, h->A);
    h->launches += 1;
    return RTS_OK;
}

int32_t rts_set_agents(RtsHandle h, uint32_t n, const RtsAgentView* agents) {
    if (!h) return RTS_E_ARG;
    if (n && (!agents || !agents->r)) return h->fail(RTS_E_ARG, "rts_set_agents", "positions are required");
    if (n >= (1u << 28)) return h->fail(RTS_E_ARG, "rts_set_agents", "more than 2^28 agents");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    CU(cudaStreamSynchronize(h->stream));
    int32_t rc;
    if (h->multi) {
        if (n && !agents->gid) return h->fail(RTS_E_ARG, "rts_set_agents", "a slab handle needs global ids (RtsAgentView.gid)");
        const uint32_t extra = 2 * (h->SL.cap_mig + h->SL.cap_halo);
        const uint32_t want_main = n + n / 4 + 4096;
        if (want_main > h->cap_main || h->cap < h->cap_main + extra) {
            h->cap_main = want_main;
            if ((rc = alloc_agents(h, h->cap_main + extra)) != RTS_OK) return rc;
        }
    } else if (n > h->cap || h->cap == 0) {
        const uint32_t cap = std::max<uint32_t>(n + n / 8 + 1024, 1024);
        if ((rc = alloc_agents(h, cap)) != RTS_OK) return rc;
    }
    if (agents && agents->gid)
        for (uint32_t i = 0; i < n; ++i)
            if (agents->gid[i] >= (1u << 28)) return h->fail(RTS_E_ARG, "rts_set_agents", "gid >= 2^28 (the ordering key keeps 28 bits of it)");
    if (h->fine_cs_auto && n) {
        // Gather cells sized for ~0.8 agents each (never below the interaction range: setup_fine_grid clamps): on a sparse map
        // (config 4: 0.006 agents per unit²) 6-unit cells would be 80 % empty and the per-frame scan over them — 46 M cells on
        // 8192² — would cost more than the neighbour search saves.
        // Density = agents over their bounding box (clipped to the map / the slab).
        const RtsParams& P = h->P;
        const float W = P.map_nx * P.map_cell_size, y0 = P.slab_row_lo * P.ns_cell_size, y1 = std::min(P.map_ny * P.map_cell_size, P.slab_row_hi * P.ns_cell_size);
        float bx0 = W, bx1 = 0.f, by0 = y1, by1 = y0;
        for (uint32_t i = 0; i < n; ++i) {
            const float x = agents->r[2 * i], y = agents->r[2 * i + 1];
            if (x < bx0) bx0 = x;
            if (x > bx1) bx1 = x;
            if (y < by0) by0 = y;
            if (y > by1) by1 = y;
        }
        bx0 = std::max(bx0, 0.f); bx1 = std::min(bx1, W); by0 = std::max(by0, y0); by1 = std::min(by1, y1);
        const double area = std::max(1.0, (double)std::max(bx1 - bx0, 6.f) * (double)std::max(by1 - by0, 6.f));
        const float want = std::min(16.0f, std::max(6.0f, (float)std::sqrt(0.8 * area / (double)n)));
        if (std::fabs(want - h->fine_cs) > 0.15f * h->fine_cs) {
            h->fine_cs = want;
            if ((rc = setup_fine_grid(h)) != RTS_OK) return rc;
        }
    }
    h->n = n;
    h->delta_valid = false;   // (rts_download_delta reports everyone after a population change)
    h->gid_is_index = !(agents && agents->gid);
    if ((rc = ingest(h, 0, n, agents)) != RTS_OK) return rc;
    if ((rc = rebin(h)) != RTS_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

// current state -> "n" buffers in slot order, so that new agents can be appended behind them
static int32_t flatten(RtsContext* h) {
    Staging none{};
    if (h->n) k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, none, h->A);
    h->launches += 1;
    return RTS_OK;
}

int32_t rts_append_agents(RtsHandle h, uint32_t m, const RtsAgentView* agents) {
    if (!h) return RTS_E_ARG;
    if (m == 0) return RTS_OK;
    if (!agents || !agents->r) return h->fail(RTS_E_ARG, "rts_append_agents", "positions are required");
    if (h->n == 0) return rts_set_agents(h, m, agents);
    if ((uint64_t)h->n + m >= (1u << 28)) return h->fail(RTS_E_ARG, "rts_append_agents", "more than 2^28 agents");
    if (h->gid_is_index && agents->gid) return h->fail(RTS_E_ARG, "rts_append_agents", "population uses component indices as gid");
    if (!h->gid_is_index && !agents->gid) return h->fail(RTS_E_ARG, "rts_append_agents", "population uses explicit gids");
    if (agents->gid)
        for (uint32_t i = 0; i < m; ++i)
            if (agents->gid[i] >= (1u << 28)) return h->fail(RTS_E_ARG, "rts_append_agents", "gid >= 2^28 (the ordering key keeps 28 bits of it)");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    int32_t rc;
    if (h->n + m > h->cap) {
        // grow: keep the old arrays alive while the state is copied across
        CU(cudaStreamSynchronize(h->stream));
        AgentArrays old = h->A;
        std::vector<void*> old_allocs;
        old_allocs.swap(h->agent_allocs);
        const uint32_t n_old = h->n;
        if ((rc = alloc_agents(h, (h->n + m) + (h->n + m) / 4 + 1024)) != RTS_OK) { free_all(old_allocs); return rc; }
        AgentArrays fresh = h->A;
        // flatten old state into the fresh "n" buffers
        AgentArrays mix = old;
        mix.pos4n = fresh.pos4n; mix.key2n = fresh.key2n; mix.vel4n = fresh.vel4n; mix.nav4n = fresh.nav4n;
        mix.trin = fresh.trin; mix.vel_outn = fresh.vel_outn;
        Staging none{};
        k_apply<<<blocks_for(n_old, 256), 256, 0, h->stream>>>(n_old, none, mix);
        CU(cudaStreamSynchronize(h->stream));
        free_all(old_allocs);
    } else {
        if ((rc = flatten(h)) != RTS_OK) return rc;
    }
    // gid of the old agents is kept; new ones get n..n+m-1 (component indices) or their explicit gid
    if ((rc = ingest(h, h->n, m, agents)) != RTS_OK) return rc;
    h->n += m;
    h->delta_valid = false;
    if ((rc = rebin(h)) != RTS_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

int32_t rts_update_agents(RtsHandle h, const RtsAgentView* fields) {
    if (!h || !fields) return RTS_E_ARG;
    if (!h->gid_is_index) return h->fail(RTS_E_STATE, "rts_update_agents", "needs component-index gids");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    Staging S;
    int32_t rc = stage_in(h, fields, h->n, &S);
    if (rc != RTS_OK) return rc;
    if (h->n) k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, S, h->A);
    h->launches += 1;
    return rebin(h);
}

int32_t rts_update_agents_indexed(RtsHandle h, const uint32_t* indices, uint32_t m, const RtsAgentView* fields) {
    if (!h || !fields) return RTS_E_ARG;
    if (m == 0) return RTS_OK;
    if (!indices) return h->fail(RTS_E_ARG, "rts_update_agents_indexed", "null indices");
    if (!h->gid_is_index) return h->fail(RTS_E_STATE, "rts_update_agents_indexed", "needs component-index gids");
    if (m > h->n) return h->fail(RTS_E_ARG, "rts_update_agents_indexed", "more indices than agents");
    for (uint32_t k = 0; k < m; ++k)
        if (indices[k] >= h->n) return h->fail(RTS_E_ARG, "rts_update_agents_indexed", "index out of range");
    if (fields->path_id)
        for (uint32_t k = 0; k < m; ++k)
            if (fields->path_id[k] >= (int32_t)h->M.n_paths) return h->fail(RTS_E_ARG, "rts_update_agents_indexed", "path_id out of range");
    CU(cudaSetDevice(h->device));
    // (only a position edit re-bins, i.e. overwrites what a pending walk queue refers to; the in-place edits below do not)
    if (fields->r) { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    Staging S;
    int32_t rc = stage_in(h, fields, m, &S);   // m entries per field, compact
    if (rc != RTS_OK) return rc;
    // index list and the slot-of-gid map live in the generic scratch buffer: [inv: cap][idx: m]
    if ((rc = ensure_scratch(h, sizeof(uint32_t) * ((size_t)h->cap + m))) != RTS_OK) return rc;
    uint32_t* inv = static_cast<uint32_t*>(h->scratch);
    uint32_t* idx = inv + h->cap;
    CU(cudaMemcpyAsync(idx, indices, sizeof(uint32_t) * (size_t)m, cudaMemcpyHostToDevice, h->stream));
    k_slot_of_gid<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, h->A.key2, inv);
    if (fields->r) {   // positions change: current state -> "n" buffers (slot order), edit there, re-bin
        Staging none{};
        k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, none, h->A);
        k_apply_indexed<true><<<blocks_for(m, 128), 128, 0, h->stream>>>(m, idx, inv, S, h->A);
        h->launches += 3;
        return rebin(h);
    }
    k_apply_indexed<false><<<blocks_for(m, 128), 128, 0, h->stream>>>(m, idx, inv, S, h->A);
    h->launches += 2;
    CU(cudaGetLastError());
    return RTS_OK;
}

uint32_t rts_agent_count(RtsHandle h) { return h ? h->n : 0; }

int32_t rts_get_finished_area(RtsHandle h, float* out, uint32_t n) {
    if (!h || !out) return RTS_E_ARG;
    if (n > h->M.n_groups) return h->fail(RTS_E_ARG, "rts_get_finished_area", "more groups requested than uploaded");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (n && !h->multi) CU(cudaMemcpy(out, h->M.area_acc, sizeof(float) * n, cudaMemcpyDeviceToHost));
    if (n && h->multi) {   // the area the next frame will see: area_read + the stops (of all ranks) not yet added to it
        std::vector<uint32_t> tot(n), done(n);
        CU(cudaMemcpy(out, h->M.area_read, sizeof(float) * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(tot.data(), h->M.stop_global, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(done.data(), h->M.stop_applied, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        const float unit = 3.14159265358979323846f * h->P.rhard * h->P.rhard;
        for (uint32_t g = 0; g < n; ++g)
            for (uint32_t k = done[g]; k < tot[g]; ++k) out[g] += unit;
    }
    return RTS_OK;
}

int32_t rts_set_finished_area(RtsHandle h, const float* area, uint32_t n) {
    if (!h || !area) return RTS_E_ARG;
    if (n > h->M.n_groups) return h->fail(RTS_E_ARG, "rts_set_finished_area", "more groups than uploaded");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (n) {
        CU(cudaMemcpy(h->M.area_acc, area, sizeof(float) * n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->M.area_read, area, sizeof(float) * n, cudaMemcpyHostToDevice));
        if (h->multi) CU(cudaMemcpy(h->M.stop_applied, h->M.stop_global, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice));   // (nothing pending)
    }
    return RTS_OK;
}

constexpr uint32_t GRAPH_FRAMES = 8;   // even (the previous-order buffers ping-pong); more frames = more walks overlapped

// frames = GRAPH_FRAMES, or 1: the frame of an rts_step(h, 1) call (single-domain handles)
static int32_t build_graph(RtsContext* h, uint32_t frames = GRAPH_FRAMES) {
    const int par = h->parity;
    cudaGraphExec_t& slot = frames == 1 ? h->graph1[par] : h->graph2[par];
    if (slot) { cudaGraphExecDestroy(slot); slot = nullptr; }
    int32_t rc = join_walk(h);
    if (rc != RTS_OK) return rc;
    cudaGraph_t g = nullptr;
    const uint64_t launches0 = h->launches;
    // The first frame kernel of the graph searches the walk queue left by the frame before the graph (on a replay: by the
    // last frame of the previous replay); that queue always has the parity opposite to the graph's first frame. Searching
    // a queue that join_walk has already emptied is harmless (same positions, same answers).
    const bool deferred_walk = h->P.walk_all != 0 && h->overlap_walk && !h->profiling;
    if (deferred_walk) { h->walkq_pending = true; h->walkq_parity = par ^ 1; }
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (uint32_t f = 0; f < frames && rc == RTS_OK; ++f) rc = launch_frame(h);
    h->walkq_pending = false;   // (nothing has run; rts_step marks the queue pending after every replay)
    cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
    if (frames & 1u) swap_prev_order(h);   // (the capture swapped the host's view of the ping-pong buffers an odd number of times)
    (frames == 1 ? h->graph1_launches : h->graph_launches) = h->launches - launches0;
    h->launches = launches0;
    if (rc != RTS_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess) return h->fail(RTS_E_CUDA, "cudaStreamEndCapture", cudaGetErrorString(ce));
    CU(cudaGraphInstantiate(&slot, g, 0));
    cudaGraphDestroy(g);
    (frames == 1 ? h->graph1_n[par] : h->graph_n[par]) = h->n;
    return RTS_OK;
}

int32_t rts_step(RtsHandle h, uint32_t n_steps) {
    if (!h) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_step", "rts_upload_map has not been called");
    CU(cudaSetDevice(h->device));
    uint32_t done = 0;
    bool ev0_recorded = false;
    // (slab handles: only with the peer exchange — every node is a kernel; capturing ncclSend/ncclRecv measured slower than
    //  launching frame by frame)
    // (... and not with the arrival rule on a slab handle: its frames carry an ncclAllReduce)
    if (h->use_graph && !h->profiling && (!h->multi || ((h->graph_slabs || h->SL.peer_mode) && !h->P.enable_arrival)) && n_steps >= GRAPH_FRAMES) {
        const int par = h->parity;
        // (a slab handle's kernels take the live count from the device and cover fixed slot capacities: its graph does not
        //  depend on h->n, which rts_sync refreshes with a different value after every frame)
        if (!h->graph2[par] || (!h->multi && h->graph_n[par] != h->n)) {
            int32_t rc = build_graph(h);
            if (rc != RTS_OK) return rc;
        }
        const bool deferred_walk = h->P.walk_all != 0 && h->overlap_walk;
        if (!deferred_walk) {
            const int32_t rc = join_walk(h);
            if (rc != RTS_OK) return rc;
        }
        CU(cudaEventRecord(h->ev0, h->stream));   // (after a possible graph build: capture is host work, the device idles)
        ev0_recorded = true;
        for (; done + GRAPH_FRAMES <= n_steps; done += GRAPH_FRAMES) {
            CU(cudaGraphLaunch(h->graph2[par], h->stream));
            h->launches += h->graph_launches;
            if (deferred_walk) { h->walkq_pending = true; h->walkq_parity = par ^ 1; }   // the last frame's queue
        }
    }
    if (!ev0_recorded) CU(cudaEventRecord(h->ev0, h->stream));
    // the remaining frames — every frame of the adapter's rts_step(h, 1) — one by one: single-domain handles replay a one-frame
    // graph per ping-pong parity (the host's view of the buffers is swapped by hand, as launch_frame would have)
    const bool one_frame_graphs = h->use_graph && !h->profiling && !h->multi && h->n > 0;
    for (; done < n_steps; ++done) {
        if (one_frame_graphs) {
            const int par = h->parity;
            if (!h->graph1[par] || h->graph1_n[par] != h->n) {
                const int32_t rc = build_graph(h, 1u);
                if (rc != RTS_OK) return rc;
            }
            const bool deferred_walk = h->P.walk_all != 0 && h->overlap_walk;
            if (!deferred_walk) {
                const int32_t rc = join_walk(h);
                if (rc != RTS_OK) return rc;
            }
            CU(cudaGraphLaunch(h->graph1[par], h->stream));
            swap_prev_order(h);
            h->launches += h->graph1_launches;
            if (deferred_walk) { h->walkq_pending = true; h->walkq_parity = par; }   // this frame's queue
            continue;
        }
        const int32_t rc = launch_frame(h);
        if (rc != RTS_OK) return rc;
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaGetLastError());
    h->steps += n_steps;
    return RTS_OK;
}

// join = false: a pending walk queue is left to the leading CTAs of the next frame kernel (the caller reads no triangle index)
static int32_t sync_impl(RtsHandle h, bool join) {
    if (!h) return RTS_E_ARG;
    CU(cudaSetDevice(h->device));
    if (join) {
        const int32_t rcj = join_walk(h);
        if (rcj != RTS_OK) return rcj;
    }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    timer_collect(h);
    if (h->multi) {
        uint32_t n_live = 0, overflow = 0;
        CU(cudaMemcpy(&n_live, h->n_dev, 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&overflow, h->SL.overflow, 4, cudaMemcpyDeviceToHost));
        h->n = n_live;
        if (overflow == 3u) return h->fail(RTS_E_NCCL, "slab exchange", "the neighbour's records never arrived (peer exchange timed out after 20 s: a rank died or fell out of step)");
        if (overflow == 2u) return h->fail(RTS_E_STATE, "slab exchange", "an agent outside the edge launch packed for a neighbour (more agents in the edge rows than the halo capacity allows for, or a move longer than its unit type allows)");
        if (overflow) return h->fail(RTS_E_CAPACITY, "slab exchange", "migrant / halo buffer overflow (raise the capacities of rts_slab_configure)");
        if (n_live > h->cap_main) return h->fail(RTS_E_CAPACITY, "slab", "population outgrew the slot capacity of this handle");
    }
    return RTS_OK;
}
int32_t rts_sync(RtsHandle h) { return sync_impl(h, true); }

int32_t rts_download(RtsHandle h, const RtsAgentView* out) {
    if (!h || !out) return RTS_E_ARG;
    CU(cudaSetDevice(h->device));
    // the triangle indices of the last frame are complete only after its walk queue has been searched; a download that does
    // not ask for them (the per-frame communicate of the System2 adapter: vel, angle_vel, state, target) leaves the queue to
    // the next frame kernel, as between two frames of one rts_step call
    const bool need_walk = out->tri_idx != nullptr || out->flags != nullptr;
    if (need_walk) { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    int32_t rc = alloc_staging(h, h->cap);
    if (rc != RTS_OK) return rc;
    Staging S = h->S;
#define WANT(field) if (!out->field) S.field = nullptr;
    WANT(r) WANT(vel) WANT(inertia) WANT(angle) WANT(angle_vel) WANT(radius) WANT(mass) WANT(state) WANT(player)
    WANT(unit_type) WANT(path_id) WANT(waypoint) WANT(r_next) WANT(gid) WANT(tri_idx) WANT(ns_cell) WANT(map_cell) WANT(flags)
#undef WANT
    if (h->multi) {
        rc = sync_impl(h, need_walk);
        if (rc != RTS_OK) return rc;
    }
    const uint32_t n = h->n;
    if (n) k_gather<<<blocks_for(n, 256), 256, 0, h->stream>>>(n, h->P, h->A, S, h->gid_is_index ? 1 : 0, out->flags ? 1 : 0);
    h->launches += 1;
#define DOWN(field, T, k) \
    if (out->field) CU(cudaMemcpyAsync(out->field, S.field, sizeof(T) * (size_t)(k) * n, cudaMemcpyDeviceToHost, h->stream));
    DOWN(r, float, 2) DOWN(vel, float, 2) DOWN(inertia, float, 2) DOWN(angle, float, 1) DOWN(angle_vel, float, 1)
    DOWN(radius, float, 1) DOWN(mass, float, 1) DOWN(state, uint8_t, 1) DOWN(player, uint8_t, 1) DOWN(unit_type, uint8_t, 1)
    DOWN(path_id, int32_t, 1) DOWN(waypoint, int32_t, 1) DOWN(r_next, float, 2) DOWN(gid, uint32_t, 1) DOWN(tri_idx, uint32_t, 1)
    DOWN(ns_cell, uint32_t, 1) DOWN(map_cell, uint32_t, 1) DOWN(flags, uint32_t, 1)
#undef DOWN
    return sync_impl(h, need_walk);
}

int32_t rts_download_delta(RtsHandle h, float* vel, RtsAgentDelta* changes, uint32_t cap, uint32_t* n_changes,
                           float* angle_vel, uint8_t* state, float* target) {
    if (!h || !vel || !n_changes || (cap && !changes)) return RTS_E_ARG;
    static_assert(sizeof(DeltaRec) == sizeof(RtsAgentDelta), "device record = ABI record");
    *n_changes = 0;
    if (h->multi || !h->gid_is_index) return h->fail(RTS_E_STATE, "rts_download_delta", "needs a single-domain handle with component-index gids");
    CU(cudaSetDevice(h->device));
    const uint32_t n = h->n;
    if (n == 0) return sync_impl(h, false);
    int32_t rc = alloc_staging(h, h->cap);
    if (rc != RTS_OK) return rc;
    if (h->delta_cap < h->cap) {
        if (h->delta_shadow) cudaFree(h->delta_shadow);
        if (h->delta_changes) cudaFree(h->delta_changes);
        h->delta_shadow = nullptr; h->delta_changes = nullptr; h->delta_cap = 0;
        CU(cudaMalloc(&h->delta_shadow, sizeof(float4) * (size_t)h->cap));
        CU(cudaMalloc(&h->delta_changes, sizeof(DeltaRec) * (size_t)h->cap));
        if (!h->delta_count) CU(cudaMalloc(&h->delta_count, 16));
        if (!h->delta_count_host) CU(cudaHostAlloc(&h->delta_count_host, 16, cudaHostAllocDefault));
        if (!h->ev_delta) CU(cudaEventCreateWithFlags(&h->ev_delta, cudaEventDisableTiming));
        h->delta_cap = h->cap;
        h->delta_valid = false;
    }
    CU(cudaMemsetAsync(h->delta_count, 0, 4, h->stream));
    k_gather_delta<<<blocks_for(n, 256), 256, 0, h->stream>>>(n, h->A, h->S.vel, h->delta_shadow, h->delta_changes, h->delta_count, h->delta_cap, h->delta_valid ? 0 : 1);
    h->launches += 1;
    // the count first (the host needs it to size the next copy), then the change list, then the velocities — the big copy last,
    // so that the host can write the changes into the blackboard while the velocities are still arriving
    CU(cudaMemcpyAsync(h->delta_count_host, h->delta_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev_delta, h->stream));
    CU(cudaEventSynchronize(h->ev_delta));
    const uint32_t count = *h->delta_count_host;
    h->delta_valid = true;
    if (count > cap) {   // (the shadow has moved on: start over with the next call)
        h->delta_valid = false;
        sync_impl(h, false);
        return h->fail(RTS_E_CAPACITY, "rts_download_delta", "more agents changed than the change list holds (cap = rts_agent_count() always suffices)");
    }
    if (count) CU(cudaMemcpyAsync(changes, h->delta_changes, sizeof(DeltaRec) * (size_t)count, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev_delta, h->stream));
    CU(cudaMemcpyAsync(vel, h->S.vel, sizeof(float) * 2 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    *n_changes = count;
    if (count && (angle_vel || state || target)) {
        CU(cudaEventSynchronize(h->ev_delta));
        const int32_t rca = rts_apply_delta(changes, count, n, angle_vel, state, target);
        if (rca != RTS_OK) { sync_impl(h, false); return h->fail(rca, "rts_download_delta", "change list entry out of range"); }
    }
    return sync_impl(h, false);
}

int32_t rts_apply_delta(const RtsAgentDelta* changes, uint32_t n_changes, uint32_t n, float* angle_vel, uint8_t* state, float* target) {
    if (n_changes && !changes) return RTS_E_ARG;
    constexpr uint32_t AHEAD = 24;   // the entries come in the device's cell order, i.e. at random places of the arrays: the
                                     // lines of the entries a little further down the list are requested early
    for (uint32_t k = 0; k < n_changes; ++k) {
        if (k + AHEAD < n_changes) {
            const uint32_t j = changes[k + AHEAD].index;
            if (j < n) {
                if (angle_vel) __builtin_prefetch(angle_vel + j, 1);
                if (state) __builtin_prefetch(state + j, 1);
                if (target) __builtin_prefetch(target + 2 * (size_t)j, 1);
            }
        }
        const RtsAgentDelta& c = changes[k];
        if (c.index >= n) return RTS_E_ARG;
        if (angle_vel) angle_vel[c.index] = c.angle_vel;
        if (state) state[c.index] = (uint8_t)c.state;
        if (target) { target[2 * (size_t)c.index] = c.target_x; target[2 * (size_t)c.index + 1] = c.target_y; }
    }
    return RTS_OK;
}

// swap-with-last removal (System2::remove, ECS.h:197-214): rare, done through the host
int32_t rts_remove_agents(RtsHandle h, const uint32_t* indices, uint32_t m) {
    if (!h) return RTS_E_ARG;
    if (m == 0) return RTS_OK;
    if (!indices) return h->fail(RTS_E_ARG, "rts_remove_agents", "null indices");
    if (!h->gid_is_index) return h->fail(RTS_E_STATE, "rts_remove_agents", "needs component-index gids");
    uint32_t n = h->n;
    std::vector<float> r(2 * (size_t)n), inertia(2 * (size_t)n), angle(n), angle_vel(n), radius(n), mass(n), r_next(2 * (size_t)n);
    std::vector<uint8_t> state(n), player(n), unit_type(n);
    std::vector<int32_t> path_id(n), waypoint(n);
    RtsAgentView v{};
    v.r = r.data(); v.inertia = inertia.data(); v.angle = angle.data(); v.angle_vel = angle_vel.data(); v.radius = radius.data();
    v.mass = mass.data(); v.state = state.data(); v.player = player.data(); v.unit_type = unit_type.data();
    v.path_id = path_id.data(); v.waypoint = waypoint.data(); v.r_next = r_next.data();
    int32_t rc = rts_download(h, &v);
    if (rc != RTS_OK) return rc;
    for (uint32_t k = 0; k < m; ++k) {
        const uint32_t i = indices[k];
        if (i >= n) return h->fail(RTS_E_ARG, "rts_remove_agents", "index out of range");
        const uint32_t l = n - 1;
        r[2 * i] = r[2 * l]; r[2 * i + 1] = r[2 * l + 1];
        inertia[2 * i] = inertia[2 * l]; inertia[2 * i + 1] = inertia[2 * l + 1];
        r_next[2 * i] = r_next[2 * l]; r_next[2 * i + 1] = r_next[2 * l + 1];
        angle[i] = angle[l]; angle_vel[i] = angle_vel[l]; radius[i] = radius[l]; mass[i] = mass[l];
        state[i] = state[l]; player[i] = player[l]; unit_type[i] = unit_type[l]; path_id[i] = path_id[l]; waypoint[i] = waypoint[l];
        --n;
    }
    return rts_set_agents(h, n, &v);
}

// ECSystem::setMoveState + SeekSystem::issuePaths2 for m agents (Game.cpp:211-231)
int32_t rts_command_agents(RtsHandle h, const uint32_t* indices, uint32_t m, int32_t path_id, uint8_t state) {
    if (!h) return RTS_E_ARG;
    if (m == 0) return RTS_OK;
    if (!indices) return h->fail(RTS_E_ARG, "rts_command_agents", "null indices");
    if (!h->gid_is_index) return h->fail(RTS_E_STATE, "rts_command_agents", "needs component-index gids");
    if (path_id >= (int32_t)h->M.n_paths) return h->fail(RTS_E_ARG, "rts_command_agents", "path_id out of range");
    // only the commanded agents travel (rts_update_agents_indexed): state, path, waypoint 0, r_next = "take it from the path"
    std::vector<uint8_t> st(m, state);
    std::vector<int32_t> pid(m, path_id), wp(m, 0);
    std::vector<float> rn(2 * (size_t)m, std::nanf(""));
    RtsAgentView v{};
    v.state = st.data(); v.path_id = pid.data(); v.waypoint = wp.data(); v.r_next = rn.data();
    const int32_t rc = rts_update_agents_indexed(h, indices, m, &v);
    if (rc != RTS_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));   // the staging copies read these host vectors
    return RTS_OK;
}

int32_t rts_get_timings(RtsHandle h, RtsTimings* out) {
    if (!h || !out) return RTS_E_ARG;
    std::memset(out, 0, sizeof(*out));
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    timer_collect(h);
    float ms = 0;
    if (h->steps && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) out->last_step_ms = ms;
    for (int k = 0; k < KT_COUNT; ++k) out->kernel_ms[k] = h->kt[k].total_ms;
    out->kernel_launches = h->launches;
    out->steps = h->steps;
    return RTS_OK;
}

int32_t rts_kernel_time_ms(RtsHandle h, int32_t which, double* avg_ms, uint64_t* launches) {
    if (!h || which < 0 || which >= KT_COUNT) return RTS_E_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    timer_collect(h);
    const KernelTimer& t = h->kt[which];
    if (avg_ms) *avg_ms = t.launches ? t.total_ms / (double)t.launches : 0.0;
    if (launches) *launches = t.launches;
    return RTS_OK;
}

int32_t rts_reset_timings(RtsHandle h) {
    if (!h) return RTS_E_ARG;
    for (auto& t : h->kt) { t.total_ms = 0; t.launches = 0; t.used = 0; }
    h->launches = 0;
    h->steps = 0;
    return RTS_OK;
}

int32_t rts_set_profiling(RtsHandle h, int32_t on) {
    if (!h) return RTS_E_ARG;
    h->profiling = on != 0;
    return RTS_OK;
}

// ---- stand-alone queries ------------------------------------------------------------------------
int32_t rts_coord_to_cell(RtsHandle h, const float* xy, uint32_t n, int32_t nx, int32_t ny, float cs, uint32_t* out) {
    if (!h) return RTS_E_ARG;
    if (n == 0) return RTS_OK;
    if (!xy || !out || nx <= 0 || ny <= 0 || !(cs > 0)) return h->fail(RTS_E_ARG, "rts_coord_to_cell", "bad argument");
    CU(cudaSetDevice(h->device));
    int32_t rc = ensure_scratch(h, (size_t)n * 12);
    if (rc != RTS_OK) return rc;
    float2* d_xy = static_cast<float2*>(h->scratch);
    uint32_t* d_out = reinterpret_cast<uint32_t*>(d_xy + n);
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    k_coord_to_cell<<<blocks_for(n, 256), 256, 0, h->stream>>>(d_xy, n, nx, ny, cs, d_out);
    CU(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

int32_t rts_find_triangles(RtsHandle h, const float* xy, uint32_t n, uint32_t* out) {
    if (!h) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_find_triangles", "rts_upload_map has not been called");
    if (n == 0) return RTS_OK;
    if (!xy || !out) return h->fail(RTS_E_ARG, "rts_find_triangles", "null");
    CU(cudaSetDevice(h->device));
    int32_t rc = ensure_scratch(h, (size_t)n * 12);
    if (rc != RTS_OK) return rc;
    float2* d_xy = static_cast<float2*>(h->scratch);
    uint32_t* d_out = reinterpret_cast<uint32_t*>(d_xy + n);
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    k_find_triangles<<<blocks_for(n, 128), 128, 0, h->stream>>>(d_xy, n, h->P, h->M, d_out);
    CU(cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

int32_t rts_build_cell_grid(RtsHandle h, uint32_t last_found_tri, uint32_t* cell2tri_out) {
    if (!h) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_build_cell_grid", "rts_upload_map has not been called");
    CU(cudaSetDevice(h->device));
    int32_t rc = join_walk(h);   // a walk of the previous frame may still be reading the cache
    if (rc != RTS_OK) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if ((rc = build_cell_grid(h, last_found_tri)) != RTS_OK) return rc;
    if ((rc = build_border_cache(h)) != RTS_OK) return rc;
    h->drop_graphs();
    if (cell2tri_out)
        CU(cudaMemcpy(cell2tri_out, h->c2t_dev, sizeof(uint32_t) * (size_t)h->P.tri_nx * h->P.tri_ny, cudaMemcpyDeviceToHost));
    return RTS_OK;
}

int32_t rts_contact_edges(RtsHandle h, const float* xy, const float* radius, uint32_t n, uint32_t cap, uint32_t* counts, float* edges_out) {
    if (!h) return RTS_E_ARG;
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_contact_edges", "rts_upload_map has not been called");
    if (n == 0) return RTS_OK;
    if (!xy || !radius || !counts) return h->fail(RTS_E_ARG, "rts_contact_edges", "null");
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)n * 16 + (edges_out ? (size_t)n * cap * 20 : 0);
    int32_t rc = ensure_scratch(h, bytes);
    if (rc != RTS_OK) return rc;
    float2* d_xy = static_cast<float2*>(h->scratch);
    float* d_rad = reinterpret_cast<float*>(d_xy + n);
    uint32_t* d_cnt = reinterpret_cast<uint32_t*>(d_rad + n);
    float* d_edges = edges_out ? reinterpret_cast<float*>(d_cnt + n) : nullptr;
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_rad, radius, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    if (d_edges) CU(cudaMemsetAsync(d_edges, 0, (size_t)n * cap * 20, h->stream));
    k_contact_edges<<<blocks_for(n, 128), 128, 0, h->stream>>>(d_xy, d_rad, n, cap, h->P, h->M, d_cnt, d_edges);
    CU(cudaMemcpyAsync(counts, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (d_edges) CU(cudaMemcpyAsync(edges_out, d_edges, (size_t)n * cap * 20, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

// ---- f4, first slice -----------------------------------------------------------------------------------------------
int32_t rts_query_radius(RtsHandle h, const float* xy, const float* r_max, uint32_t n, uint32_t cap, uint32_t* counts, uint32_t* indices_out) {
    if (!h) return RTS_E_ARG;
    if (n == 0) return RTS_OK;
    if (!xy || !r_max || !counts || (cap && !indices_out)) return h->fail(RTS_E_ARG, "rts_query_radius", "null");
    if (h->multi) return h->fail(RTS_E_STATE, "rts_query_radius", "single-domain handles only");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    const size_t bytes = (size_t)n * 16 + (size_t)n * cap * 8;
    int32_t rc = ensure_scratch(h, bytes);
    if (rc != RTS_OK) return rc;
    float2* d_xy = static_cast<float2*>(h->scratch);
    float* d_r = reinterpret_cast<float*>(d_xy + n);
    uint32_t* d_cnt = reinterpret_cast<uint32_t*>(d_r + n);
    uint2* d_out = reinterpret_cast<uint2*>(d_cnt + n);
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(d_r, r_max, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    k_query_radius<<<blocks_for(n, 64), 64, 0, h->stream>>>(n, d_xy, d_r, cap, h->P, h->F, h->A, d_cnt, d_out);
    h->launches += 1;
    std::vector<uint2> pairs((size_t)n * cap);
    CU(cudaMemcpyAsync(counts, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (cap) CU(cudaMemcpyAsync(pairs.data(), d_out, pairs.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    for (uint32_t q = 0; q < n; ++q) {   // the reference's visit order: cells row-major, ascending component index inside a cell
        uint2* first = pairs.data() + (size_t)q * cap;
        const uint32_t m = std::min(counts[q], cap);
        std::sort(first, first + m, [](const uint2& a, const uint2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
        for (uint32_t k = 0; k < m; ++k) indices_out[(size_t)q * cap + k] = first[k].y;
    }
    return RTS_OK;
}

int32_t rts_attack_targets(RtsHandle h, float range2, int32_t* out) {
    if (!h || !out) return RTS_E_ARG;
    if (h->multi) return h->fail(RTS_E_STATE, "rts_attack_targets", "single-domain handles only");
    if (!(range2 > 0)) return h->fail(RTS_E_ARG, "rts_attack_targets", "range2 must be positive");
    if (h->n == 0) return RTS_OK;
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    int32_t rc = ensure_scratch(h, (size_t)h->n * 4);
    if (rc != RTS_OK) return rc;
    int32_t* d_out = static_cast<int32_t*>(h->scratch);
    k_attack_targets<<<blocks_for(h->n, 128), 128, 0, h->stream>>>(h->n, range2, h->P, h->F, h->A, d_out, h->gid_is_index ? 1 : 0);
    h->launches += 1;
    CU(cudaMemcpyAsync(out, d_out, (size_t)h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

// ---- slab decomposition ---------------------------------------------------------------------------
int32_t rts_slab_configure(RtsHandle h, int32_t rank, int32_t nranks, uint32_t cap_migrants, uint32_t cap_halo) {
    if (!h) return RTS_E_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return h->fail(RTS_E_ARG, "rts_slab_configure", "bad rank");
    if (h->n) return h->fail(RTS_E_STATE, "rts_slab_configure", "configure the slab before uploading agents");
    if (h->P.enable_alignment) return h->fail(RTS_E_STATE, "rts_slab_configure", "enable_alignment is not supported on a slab handle (halo records carry no velocity)");
    if (h->P.enable_arrival && !(h->P.finish_radius + 0.25f < h->P.ns_cell_size))
        return h->fail(RTS_E_ARG, "rts_slab_configure", "enable_arrival on a slab handle needs finish_radius < ns_cell_size");
    const RtsParams& P = h->P;
    if (P.slab_row_lo < 0 || P.slab_row_hi > P.ns_ny || P.slab_row_lo >= P.slab_row_hi) return h->fail(RTS_E_ARG, "rts_slab_configure", "bad slab rows in the params");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    free_all(h->slab_allocs);
    cap_migrants = (cap_migrants + 3u) & ~3u;
    cap_halo = (cap_halo + 3u) & ~3u;
    if (cap_migrants >= (1u << 20) || cap_halo >= (1u << 20)) return h->fail(RTS_E_ARG, "rts_slab_configure", "exchange capacities must stay below 2^20 records");
    SlabDev SL{};
    SL.row_lo = P.slab_row_lo; SL.row_hi = P.slab_row_hi;
    SL.y_lo = P.slab_row_lo * P.ns_cell_size; SL.y_hi = P.slab_row_hi * P.ns_cell_size;
    SL.halo = halo_band(P);
    SL.has[0] = P.slab_row_lo > 0; SL.has[1] = P.slab_row_hi < P.ns_ny;
    SL.cap_mig = cap_migrants; SL.cap_halo = cap_halo;
    const size_t bytes = exbuf_bytes(cap_migrants, cap_halo);   // layout: carve_exbuf
    auto carve = [&](void* base, ExBuf& B) { B = carve_exbuf(base, cap_migrants, cap_halo); };
    for (int d = 0; d < 2; ++d) {
        char* a = nullptr; char* b = nullptr;
        int32_t rc;
        if ((rc = dev_alloc<char>(h, &a, bytes, h->slab_allocs)) != RTS_OK) return rc;
        if ((rc = dev_alloc<char>(h, &b, bytes, h->slab_allocs)) != RTS_OK) return rc;
        CU(cudaMemsetAsync(a, 0, bytes, h->stream));
        CU(cudaMemsetAsync(b, 0, bytes, h->stream));
        h->ex_send[d] = a; h->ex_recv[d] = b;
        carve(a, SL.send[d]);
        carve(b, SL.recv[d]);
    }
    {
        uint32_t* ov = nullptr;
        int32_t rc = dev_alloc<uint32_t>(h, &ov, 4, h->slab_allocs);
        if (rc != RTS_OK) return rc;
        CU(cudaMemsetAsync(ov, 0, 16, h->stream));
        SL.overflow = ov;
    }
    h->ex_bytes = bytes;
    h->SL = SL;
    h->multi = true;
    h->rank = rank; h->nranks = nranks;
    h->drop_graphs();
    CU(cudaStreamSynchronize(h->stream));
    return setup_fine_grid(h);
}

// Peer exchange set-up (all ranks, collective): every rank allocates its receive arena, the CUDA IPC handles travel through
// ncclAllGather, each rank maps its neighbours' arenas. All or nothing: unless EVERY rank succeeded (ncclAllReduce of a
// flag) the handle stays on the ncclSend/ncclRecv exchange.
static int32_t setup_peer_exchange(RtsContext* h) {
    SlabDev& SL = h->SL;
    SL.peer_mode = 0;
    if (h->nranks < 2) return RTS_OK;
    const size_t buf = exbuf_bytes(SL.cap_mig, SL.cap_halo);
    const size_t flags_off = 4 * buf, epoch_off = flags_off + 64, arena_bytes = epoch_off + 64;
    int ok = h->want_peer ? 1 : 0;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && (cudaMalloc(&h->arena, arena_bytes) != cudaSuccess || cudaMemset(h->arena, 0, arena_bytes) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine, h->arena) != cudaSuccess)) ok = 0;
    cudaGetLastError();
    // handles of all ranks
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    char *d_mine = nullptr, *d_all = nullptr;
    int* d_ok = nullptr;
    CU(cudaMalloc(&d_mine, 64));
    CU(cudaMalloc(&d_all, 64 * (size_t)h->nranks));
    CU(cudaMalloc(&d_ok, sizeof(int)));
    CU(cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = h->p_ncclAllGather(d_mine, d_all, 64, ncclUint8, h->comm, h->stream);
    if (r != ncclSuccess) return h->fail(RTS_E_NCCL, "ncclAllGather", h->p_ncclGetErrorString(r));
    std::vector<cudaIpcMemHandle_t> all((size_t)h->nranks);
    CU(cudaMemcpyAsync(all.data(), d_all, 64 * (size_t)h->nranks, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int d = 0; d < 2 && ok; ++d) {
        if (!SL.has[d]) continue;
        const int peer = h->rank + (d == 0 ? -1 : 1);
        if (cudaIpcOpenMemHandle(&h->peer_arena[d], all[(size_t)peer], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    r = h->p_ncclAllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm, h->stream);
    if (r != ncclSuccess) return h->fail(RTS_E_NCCL, "ncclAllReduce", h->p_ncclGetErrorString(r));
    CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_mine); cudaFree(d_all); cudaFree(d_ok);
    if (!ok) {   // somebody could not: everybody keeps NCCL
        for (auto& pa : h->peer_arena)
            if (pa) { cudaIpcCloseMemHandle(pa); pa = nullptr; }
        if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
        return RTS_OK;
    }
    char* base = static_cast<char*>(h->arena);
    for (int d = 0; d < 2; ++d) {
        for (int par = 0; par < 2; ++par) {
            SL.recv_base[d][par] = base + (size_t)(2 * d + par) * buf;
            SL.peer_base[d][par] = h->peer_arena[d] ? static_cast<char*>(h->peer_arena[d]) + (size_t)(2 * (1 - d) + par) * buf : nullptr;
        }
        SL.flag_local[d] = reinterpret_cast<unsigned long long*>(base + flags_off) + d;
        SL.flag_peer[d] = h->peer_arena[d] ? reinterpret_cast<unsigned long long*>(static_cast<char*>(h->peer_arena[d]) + flags_off) + (1 - d) : nullptr;
    }
    SL.epoch = reinterpret_cast<uint32_t*>(base + epoch_off);
    SL.peer_mode = 1;
    h->drop_graphs();
    return RTS_OK;
}

int32_t rts_slab_exchange_mode(RtsHandle h) { return h ? (h->multi ? (h->SL.peer_mode ? 2 : 1) : 0) : -1; }

// ---- slab rebalancing ---------------------------------------------------------------------------------------------
int32_t rts_slab_row_counts(RtsHandle h, uint32_t* counts) {
    if (!h || !counts) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_slab_row_counts", "not a slab handle");
    CU(cudaSetDevice(h->device));
    const uint32_t rows = (uint32_t)h->P.ns_ny;
    int32_t rc = ensure_scratch(h, sizeof(uint32_t) * (size_t)rows);
    if (rc != RTS_OK) return rc;
    uint32_t* hist = static_cast<uint32_t*>(h->scratch);
    CU(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)rows, h->stream));
    const uint32_t n_slots = h->cap_main + 2 * (h->SL.cap_mig + h->SL.cap_halo);
    if (n_slots) k_row_hist<<<blocks_for(n_slots, 256), 256, 0, h->stream>>>(h->n_dev, h->A.key2, hist, rows);
    h->launches += 1;
    CU(cudaMemcpyAsync(counts, hist, sizeof(uint32_t) * (size_t)rows, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return RTS_OK;
}

int32_t rts_slab_set_rows(RtsHandle h, int32_t row_lo, int32_t row_hi) {
    if (!h) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_slab_set_rows", "not a slab handle");
    if (row_lo < 0 || row_hi > h->P.ns_ny || row_lo >= row_hi) return h->fail(RTS_E_ARG, "rts_slab_set_rows", "bad rows");
    if ((row_lo > 0) != (h->SL.has[0] != 0) || (row_hi < h->P.ns_ny) != (h->SL.has[1] != 0))
        return h->fail(RTS_E_ARG, "rts_slab_set_rows", "the outermost slabs keep the map edge: only boundaries shared with a neighbour move");
    if (row_lo == h->P.slab_row_lo && row_hi == h->P.slab_row_hi) return RTS_OK;
    CU(cudaSetDevice(h->device));
    int32_t rc = rts_sync(h);   // also refreshes h->n (owned + ghost slots)
    if (rc != RTS_OK) return rc;
    h->P.slab_row_lo = row_lo; h->P.slab_row_hi = row_hi;
    SlabDev& SL = h->SL;
    SL.row_lo = row_lo; SL.row_hi = row_hi;
    SL.y_lo = row_lo * h->P.ns_cell_size; SL.y_hi = row_hi * h->P.ns_cell_size;
    // the gather grid follows the slab (one physics row of margin on either side): current state -> "n" buffers -> new grid.
    // Ownership is decided by the frame kernel from the physics row of each agent's position, so the agents of the rows
    // that changed hands simply migrate in the next frame (their count must fit the migrant capacity).
    if ((rc = setup_fine_grid(h)) != RTS_OK) return rc;
    if (h->n) {
        Staging none{};
        k_apply<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->n, none, h->A);
        h->launches += 1;
        if ((rc = rebin(h)) != RTS_OK) return rc;
    }
    h->drop_graphs();
    CU(cudaStreamSynchronize(h->stream));
    return RTS_OK;
}

int32_t rts_comm_unique_id(void* out128) {
    if (!out128) return RTS_E_ARG;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return RTS_E_NCCL;
    auto fn = reinterpret_cast<decltype(&ncclGetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    if (!fn) return RTS_E_NCCL;
    ncclUniqueId id;
    if (fn(&id) != ncclSuccess) return RTS_E_NCCL;
    std::memcpy(out128, &id, sizeof(id));
    return RTS_OK;
}

int32_t rts_comm_init(RtsHandle h, const void* id128) {
    if (!h || !id128) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_comm_init", "rts_slab_configure first");
    CU(cudaSetDevice(h->device));
    h->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h->nccl_lib) return h->fail(RTS_E_NCCL, "dlopen(libnccl.so.2)", dlerror());
#define SYM(name)                                                                              \
    h->p_##name = reinterpret_cast<decltype(&name)>(dlsym(h->nccl_lib, #name));                 \
    if (!h->p_##name) return h->fail(RTS_E_NCCL, "dlsym", #name);
    SYM(ncclCommInitRank) SYM(ncclCommDestroy) SYM(ncclSend) SYM(ncclRecv) SYM(ncclGroupStart) SYM(ncclGroupEnd) SYM(ncclGetErrorString)
    SYM(ncclAllGather) SYM(ncclAllReduce)
#undef SYM
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    const ncclResult_t r = h->p_ncclCommInitRank(&h->comm, h->nranks, id, h->rank);
    if (r != ncclSuccess) return h->fail(RTS_E_NCCL, "ncclCommInitRank", h->p_ncclGetErrorString(r));
    return setup_peer_exchange(h);
}

// first half of a frame (frame kernel + packing of migrants / halo) — the caller moves the buffers, then rts_step_end
int32_t rts_step_begin(RtsHandle h) {
    if (!h) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_step_begin", "not a slab handle");
    if (!h->map_ready) return h->fail(RTS_E_STATE, "rts_step_begin", "rts_upload_map has not been called");
    CU(cudaSetDevice(h->device));
    const int32_t rc = frame_begin(h);
    h->steps += 1;
    return rc;
}
int32_t rts_step_end(RtsHandle h) {
    if (!h) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_step_end", "not a slab handle");
    CU(cudaSetDevice(h->device));
    return frame_end(h);
}
// packs the halo bands of the current state without stepping (after an upload); finish with rts_step_end
int32_t rts_halo_begin(RtsHandle h) {
    if (!h) return RTS_E_ARG;
    if (!h->multi) return h->fail(RTS_E_STATE, "rts_halo_begin", "not a slab handle");
    CU(cudaSetDevice(h->device));
    { const int32_t rcj = join_walk(h); if (rcj != RTS_OK) return rcj; }
    for (int d = 0; d < 2; ++d) CU(cudaMemsetAsync(h->ex_send[d], 0, 16, h->stream));
    if (h->cap_main) k_halo_seed<<<blocks_for(h->cap_main, 256), 256, 0, h->stream>>>(h->cap_main, h->P, h->F, h->A, h->SL);
    h->launches += 1;
    return RTS_OK;
}
// rts_halo_begin + NCCL exchange + rts_step_end
int32_t rts_halo_refresh(RtsHandle h) {
    int32_t rc = rts_halo_begin(h);
    if (rc != RTS_OK) return rc;
    if (h->nranks > 1 && (rc = exchange(h, false)) != RTS_OK) return rc;
    return frame_end(h);
}
// single-device emulation of two neighbouring ranks (tests): lower.send[up] -> upper.recv[down] and back
int32_t rts_slab_exchange_local(RtsHandle lower, RtsHandle upper) {
    if (!lower || !upper) return RTS_E_ARG;
    RtsHandle h = lower;
    if (!lower->multi || !upper->multi || lower->ex_bytes != upper->ex_bytes) return h->fail(RTS_E_STATE, "rts_slab_exchange_local", "handles are not matching slabs");
    CU(cudaStreamSynchronize(lower->stream));
    CU(cudaStreamSynchronize(upper->stream));
    CU(cudaStreamSynchronize(lower->stream4));   // the edge rows (which fill the send buffers) may be on the comm stream
    CU(cudaStreamSynchronize(upper->stream4));
    CU(cudaMemcpy(upper->ex_recv[0], lower->ex_send[1], lower->ex_bytes, cudaMemcpyDeviceToDevice));
    CU(cudaMemcpy(lower->ex_recv[1], upper->ex_send[0], lower->ex_bytes, cudaMemcpyDeviceToDevice));
    // ... and reduce_stop_counts over the two "ranks"
    if (lower->P.enable_arrival && upper->P.enable_arrival && lower->M.n_groups && lower->M.n_groups == upper->M.n_groups) {
        const uint32_t ng = lower->M.n_groups;
        std::vector<uint32_t> a(ng), b(ng);
        CU(cudaMemcpy(a.data(), lower->M.stop_local, sizeof(uint32_t) * ng, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(b.data(), upper->M.stop_local, sizeof(uint32_t) * ng, cudaMemcpyDeviceToHost));
        for (uint32_t g = 0; g < ng; ++g) a[g] += b[g];
        CU(cudaMemcpy(lower->M.stop_global, a.data(), sizeof(uint32_t) * ng, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(upper->M.stop_global, a.data(), sizeof(uint32_t) * ng, cudaMemcpyHostToDevice));
    }
    return RTS_OK;
}

void* rts_stream(RtsHandle h) { return h ? (void*)h->stream : nullptr; }

}  // extern "C"
