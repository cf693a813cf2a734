/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * rts_oracle.c: plain-C CPU restatement of RTSEngine's per-frame agent update, with the
 * reference's fixed-size arrays (5 000 agents, 500 per cell, 2560² box) replaced by CSR
 * cells so it runs at any size.  Same arithmetic, same operation order, same neighbour
 * visit order as the reference; built with -O2 -ffp-contract=off it is bit-identical to the
 * IEEE build of the verbatim reference (oracle/_ref/libref_harness.so) — tests/test_oracle_vs_ref.py
 * and the fixtures under tests/golden/ (made by tests/golden/make_golden.py from the verbatim
 * build) pin that.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load it; the product library never does.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "rts_gpu.h"

typedef struct { float x, y; } v2;
typedef struct { int32_t x, y; } iv2;

/* sf::Vector2 operators are component-wise (src/Utils/Vector.inl) */
static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline v2 vmul(v2 a, float s) { return V(a.x * s, a.y * s); }
static inline v2 vdiv(v2 a, float s) { return V(a.x / s, a.y / s); }
/* src/core.h:99-113 */
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vnorm2(v2 a) { return vdot(a, a); }
static inline float vnorm(v2 a) { return sqrtf(vnorm2(a)); }
static inline float vdist2(v2 a, v2 b) { return vdot(vsub(a, b), vsub(a, b)); }
static inline int vequal(v2 a, v2 b) { return vdist2(a, b) < 0.001f; }
static inline int veq(v2 a, v2 b) { return a.x == b.x && a.y == b.y; }

/* ---- Grid::coordToCell, src/Utils/Grid.cpp:18-24 -------------------------------------------
   size_t(floor(x/cs)) % nx: on x86-64 the float→size_t conversion of a negative value goes
   through a signed 64-bit convert, i.e. wraps modulo 2^64 before the %. */
static inline uint64_t f2size(float f) { return (uint64_t)(int64_t)f; }
static inline uint64_t coord_to_cell(float x, float y, int32_t nx, int32_t ny, float cs) {
    const uint64_t ix = f2size(floorf(x / cs)) % (uint64_t)nx;
    const uint64_t iy = f2size(floorf(y / cs)) % (uint64_t)ny;
    return ix + iy * (uint64_t)nx;
}

void rts_oracle_coord_to_cell(const float* xy, uint32_t n, int32_t nx, int32_t ny, float cs, uint32_t* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = (uint32_t)coord_to_cell(xy[2 * i], xy[2 * i + 1], nx, ny, cs);
}

/* ---- AcosTable<1000>, src/core.h:115-151 ---------------------------------------------------- */
#define NBINS 1000
static float g_acos[NBINS];
static int g_acos_ready = 0;
static void acos_init(void) {
    if (g_acos_ready) return;
    for (int bin = 0; bin < NBINS; ++bin) {
        const float dot_value = 2 * (float)bin / NBINS - 1;
        g_acos[bin] = 180.f / ((float)3.14159265358979323846) * acosf(dot_value);
    }
    g_acos_ready = 1;
}
void rts_oracle_acos_table(float* out) { acos_init(); memcpy(out, g_acos, sizeof(g_acos)); }

static inline float table_angle(v2 v) { /* AcosTable::angle, core.h:139-150 */
    const float dot_value = (v.x * 1.f + v.y * 0.f) / vnorm(v) + 1.0f;
    int bin = (int)floorf(dot_value / 2.f * NBINS);
    if (bin >= NBINS) bin = NBINS - 1;
    if (bin < 0) bin = 0;
    return g_acos[bin] * (2.f * (v.y > 0) - 1.f);
}
static inline float table_angle_between(v2 v1, v2 v2_) { /* AcosTable::angleBetween, core.h:126-137 */
    const float dot_value = vdot(v2_, v1) / vnorm(v1) / vnorm(v2_) + 1.0f;
    int bin = (int)floorf(dot_value / 2.f * NBINS);
    if (bin >= NBINS) bin = NBINS - 1;
    if (bin < 0) bin = 0;
    return g_acos[bin] * (2.f * ((v1.x * v2_.y - v2_.x * v1.y) > 0) - 1.f);
}

/* ---- Triangulation::findTriangle, src/PathFinding/Triangulation.cpp:1497-1567 ---------------- */
/* sign(), Triangulation.h:73-76: integer arithmetic, result converted to float.  int64 products:
   identical to the reference's int32 wherever int32 does not overflow (maps up to 1024² cells). */
static inline float tri_sign(iv2 p1, iv2 p2, iv2 p3) {
    return (float)((int64_t)(p1.x - p3.x) * (p2.y - p3.y) - (int64_t)(p2.x - p3.x) * (p1.y - p3.y));
}
static inline int in_triangle(iv2 q, const RtsTriangle* t) { /* isInTriangle, Triangulation.h:83-96 */
    const iv2 a = {t->vx[0], t->vy[0]}, b = {t->vx[1], t->vy[1]}, c = {t->vx[2], t->vy[2]};
    const float d1 = tri_sign(q, a, b), d2 = tri_sign(q, b, c), d3 = tri_sign(q, c, a);
    const int has_neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
    const int has_pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
    return !(has_neg && has_pos);
}
static inline float idot(iv2 a, iv2 b) { return (float)((int64_t)a.x * b.x + (int64_t)a.y * b.y); } /* dot<Vertex>, core.h:99 */
static inline int lines_intersect(iv2 v11, iv2 v12, iv2 v21, iv2 v22) { /* Triangulation.cpp:1805-1818 */
    const iv2 dr = {v21.x - v11.x, v21.y - v11.y};
    const iv2 t1 = {v12.x - v11.x, v12.y - v11.y};
    const iv2 t2 = {v22.x - v21.x, v22.y - v21.y};
    const iv2 n1 = {t1.y, -t1.x};
    const iv2 n2 = {t2.y, -t2.x};
    const float alpha1 = +idot(dr, n2) / idot(t1, n2);
    const float alpha2 = -idot(dr, n1) / idot(t2, n1);
    return alpha1 >= 0 && alpha1 <= 1 && alpha2 >= 0 && alpha2 <= 1;
}
static inline int nxt(int i) { return i == 2 ? 0 : i + 1; }
static inline int prv(int i) { return i == 0 ? 2 : i - 1; }

static long g_walk_brute = 0, g_walk_hops = 0, g_walk_calls = 0;
static _Thread_local uint32_t g_last_hops = 0;
void rts_oracle_walk_stats(long* out, int reset) {
    out[0] = g_walk_calls; out[1] = g_walk_hops; out[2] = g_walk_brute;
    if (reset) g_walk_calls = g_walk_hops = g_walk_brute = 0;
}

static uint32_t brute_force(iv2 q, const RtsTriangle* tris, uint32_t n_tris) {
#pragma omp atomic
    g_walk_brute++;
    for (uint32_t i = 0; i < n_tris; ++i)
        if (in_triangle(q, &tris[i])) return i;
    return 0xFFFFFFFFu;
}

/* the part of findTriangle after the start triangle is chosen (:1512-1566) */
static uint32_t walk_from(const RtsTriangle* tris, uint32_t n_tris, iv2 q, uint32_t t, uint32_t* flags) {
    if (in_triangle(q, &tris[t])) return t;
    iv2 vc;
    {
        const RtsTriangle* T = &tris[t];
        const iv2 v0 = {T->vx[0], T->vy[0]}, v1 = {T->vx[1], T->vy[1]}, v2_ = {T->vx[2], T->vy[2]};
        const iv2 g = {(v0.x + v1.x + v2_.x) / 3, (v0.y + v1.y + v2_.y) / 3}; /* integer division, :1524 */
        if (lines_intersect(g, q, v0, v1)) { t = T->neighbours[0]; vc = v0; }
        else if (lines_intersect(g, q, v1, v2_)) { t = T->neighbours[1]; vc = v1; }
        else if (lines_intersect(g, q, v2_, v0)) { t = T->neighbours[2]; vc = v2_; }
        else t = 0xFFFFFFFFu;
    }
    if (t == 0xFFFFFFFFu) return brute_force(q, tris, n_tris); /* :1539-1546 */
    uint32_t hops = 0;
    g_last_hops = 0;
    while (!in_triangle(q, &tris[t])) { /* fan walk, :1548-1563 */
        g_last_hops++;
        const RtsTriangle* T = &tris[t];
        int k = -1;
        if (T->vx[0] == vc.x && T->vy[0] == vc.y) k = 0;
        else if (T->vx[1] == vc.x && T->vy[1] == vc.y) k = 1;
        else if (T->vx[2] == vc.x && T->vy[2] == vc.y) k = 2;
        if (k < 0 || ++hops > 4 * n_tris + 16) { /* the reference would read out of bounds / spin: report */
            if (flags) *flags |= RTS_FLAG_WALK_FAILED;
            return 0xFFFFFFFFu;
        }
        const iv2 a = {T->vx[nxt(k)], T->vy[nxt(k)]}, b = {T->vx[prv(k)], T->vy[prv(k)]};
        if (lines_intersect(vc, q, a, b)) { t = T->neighbours[nxt(k)]; vc = a; }
        else t = T->neighbours[k];
        if (t == 0xFFFFFFFFu) { if (flags) *flags |= RTS_FLAG_WALK_FAILED; return t; }
    }
    return t;
}

static uint32_t find_triangle(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                              float fx, float fy, uint32_t* flags) {
    const iv2 q = {(int32_t)fx, (int32_t)fy}; /* static_cast<Vertex>(Vector2f): truncation, Triangulation.cpp:1588-1590 */
    const uint64_t cell = coord_to_cell((float)q.x, (float)q.y, P->tri_nx, P->tri_ny, P->tri_cell_size);
    const uint32_t t = cell2tri[cell];
    if (t == 0xFFFFFFFFu) return brute_force(q, tris, n_tris); /* :1503-1510 */
    return walk_from(tris, n_tris, q, t, flags);
}

/* Triangulation::updateCellGrid, Triangulation.cpp:1829-1864: the cache cells are visited in a zig-zag (row pairs
   ascending / descending, a left-over last row descending) and each cell centre is located with
   findTriangle(centre, from_last_found = true), i.e. the walk starts at the triangle found for the previous cell
   (last_found_tri_ind_, which findTriangle leaves at the returned index, :1504-1566). `last_found` is that member's value
   on entry; it matters only when the first centre lies on a triangle edge. */
void rts_oracle_build_cell_grid(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, uint32_t last_found, uint32_t* out) {
    const int nx = P->tri_nx, ny = P->tri_ny;
    const float dx = P->tri_cell_size, dy = P->tri_cell_size;
    uint32_t last = last_found;
    for (int j = 0; j < ny; ++j) {
        const int asc = (j % 2 == 0) && (j < ny - 1);
        for (int k = 0; k < nx; ++k) {
            const int i = asc ? k : nx - 1 - k;
            const float cx = i * dx + dx / 2.f, cy = j * dy + dy / 2.f;
            const iv2 q = {(int32_t)cx, (int32_t)cy};
            uint32_t t = (last == 0xFFFFFFFFu) ? brute_force(q, tris, n_tris) : walk_from(tris, n_tris, q, last, 0);
            out[(size_t)j * nx + i] = t;
            if (t != 0xFFFFFFFFu) last = t;
        }
    }
}

/* diagnostic: fan-walk hops per query point (serial) */
void rts_oracle_walk_hops(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                          const float* xy, uint32_t n, uint32_t* hops_out);
void rts_oracle_find_triangles(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                               const float* xy, uint32_t n, uint32_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = find_triangle(P, tris, n_tris, cell2tri, xy[2 * i], xy[2 * i + 1], 0);
}

/* ---- Edges::calcContactEdgesO + isTouchingEdge, src/Edges.cpp:245-296, 128-143 --------------- */
static inline int touching(const RtsEdge* e, v2 r, float radius) {
    const v2 dr = vsub(V(e->from_x, e->from_y), r);
    const v2 t = V(e->t_x, e->t_y);
    const float norm_dr_sq = vdot(dr, dr);
    const float dot_dr_t = vdot(dr, t);
    const float disc = dot_dr_t * dot_dr_t - norm_dr_sq + radius * radius;
    if (disc >= 0) {
        const float s = sqrtf(disc);
        const float a1 = -dot_dr_t + s, a2 = -dot_dr_t - s;
        return (a2 <= e->l && a2 > 0) || (a1 <= e->l && a1 > 0);
    }
    return 0;
}

typedef void (*wall_fn)(const RtsEdge* e, void* ctx);

/* visits the contact walls in the reference's order: orientation 0..3, line -D..+D, edge first..last */
static int contact_walls(const RtsParams* P, const RtsWallTable* W, v2 r, float radius, wall_fn fn, void* ctx) {
    const int ix0 = (int)(r.x / P->map_cell_size); /* Grid::cellCoords truncates, Grid.cpp:67-71 */
    const int iy0 = (int)(r.y / P->map_cell_size);
    const int base[4] = {iy0, ix0 - iy0 + P->map_ny - 1, ix0, ix0 + iy0};
    static const float dirs[4][2] = {{1.f, 0.f}, {1.f, 1.f}, {0.f, 1.f}, {1.f, -1.f}};
    const int D = (int)ceilf(radius / P->map_cell_size);
    int count = 0;
    for (int o = 0; o < 4; ++o) {
        for (int dl = -D; dl <= D; ++dl) {
            const int L = dl + base[o];
            if (L < 0 || (uint32_t)L >= W->n_lines[o]) continue;
            const uint32_t beg = W->line_offsets[o][L], end = W->line_offsets[o][L + 1];
            if (beg == end) continue;
            const RtsEdge* E = W->edges[o] + beg;
            const int size = (int)(end - beg);
            /* std::lower_bound (libstdc++ bisection) with comp(e, r) = dot(r - e.from, dir) > 0 */
            int first_i = 0, len = size;
            while (len > 0) {
                const int half = len >> 1;
                const int mid = first_i + half;
                const v2 d = vsub(r, V(E[mid].from_x, E[mid].from_y));
                if (vdot(d, V(dirs[o][0], dirs[o][1])) > 0) { first_i = mid + 1; len = len - half - 1; }
                else len = half;
            }
            int first = first_i - 1, last = first + 1;
            if (first == -1) { first = 0; last = 0; }
            if (first == size - 1) last = size - 1;
            for (int k = first; k <= last; ++k)
                if (touching(&E[k], r, radius)) { if (fn) fn(&E[k], ctx); ++count; }
        }
    }
    return count;
}

typedef struct { v2* vel; v2* inertia; } avoid_ctx;
static void avoid_one(const RtsEdge* e, void* c) { /* PhysicsSystem::avoidWall body, PhysicsSystem.cpp:146-158 */
    avoid_ctx* a = (avoid_ctx*)c;
    const v2 n = V(e->t_y, -e->t_x);
    v2 away = vmul(n, vdot(*a->vel, n));
    if (vdot(*a->vel, n) < 0) *a->vel = vsub(*a->vel, vmul(away, 1.0f));
    away = vmul(n, vdot(*a->inertia, n));
    if (vdot(*a->inertia, n) < 0) *a->inertia = vsub(*a->inertia, away);
}

typedef struct { float* out; uint32_t cap; uint32_t k; } collect_ctx;
static void collect_one(const RtsEdge* e, void* c) {
    collect_ctx* a = (collect_ctx*)c;
    if (a->k < a->cap) memcpy(a->out + 5 * a->k, e, 5 * sizeof(float));
    a->k++;
}
void rts_oracle_contact_edges(const RtsParams* P, const RtsWallTable* W, const float* xy, const float* radius, uint32_t n,
                              uint32_t cap, uint32_t* counts, float* edges_out) {
    for (uint32_t i = 0; i < n; ++i) {
        collect_ctx c = {edges_out ? edges_out + (size_t)5 * cap * i : 0, edges_out ? cap : 0, 0};
        counts[i] = (uint32_t)contact_walls(P, W, V(xy[2 * i], xy[2 * i + 1]), radius[i], collect_one, &c);
    }
}

/* ---- closestPointOnEdge, src/core.h:282-296 --------------------------------------------------- */
static inline v2 closest_point_on_edge(v2 r, const RtsEdge* e) {
    const v2 from = V(e->from_x, e->from_y), t = V(e->t_x, e->t_y);
    const v2 dr = vsub(r, from);
    const float pos = vdot(dr, t);
    if (pos < e->l && pos > 0) return vadd(from, vmul(t, pos));
    else if (pos > e->l) return vadd(from, vmul(t, e->l)); /* Edgef::to() */
    return from;
}

/* ---- the frame ------------------------------------------------------------------------------- */
typedef struct {
    uint32_t* cell_start; /* CSR over physics-grid cells */
    uint32_t* cell_items; /* agent indices, ascending within a cell (serial append, Context.cpp:31-43) */
} ns_grid;

static void build_grid(const RtsParams* P, uint32_t n, const float* r, ns_grid* g, uint32_t* cell_of) {
    const size_t nc = (size_t)P->ns_nx * P->ns_ny;
    g->cell_start = (uint32_t*)calloc(nc + 1, sizeof(uint32_t));
    g->cell_items = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) {
        cell_of[i] = (uint32_t)coord_to_cell(r[2 * i], r[2 * i + 1], P->ns_nx, P->ns_ny, P->ns_cell_size);
        g->cell_start[cell_of[i] + 1]++;
    }
    for (size_t c = 0; c < nc; ++c) g->cell_start[c + 1] += g->cell_start[c];
    uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (nc ? nc : 1));
    memcpy(fill, g->cell_start, sizeof(uint32_t) * nc);
    for (uint32_t i = 0; i < n; ++i) g->cell_items[fill[cell_of[i]]++] = i;
    free(fill);
}

static int g_slab_arrival = 0;          /* see rts_oracle_set_slab_arrival */
static uint32_t* g_stop_counts = 0;

/*
 * One frame for agents [0, n).  `ghost` (may be NULL) marks agents that only act as neighbours
 * (halo copies in the slab decomposition); they are not updated.  Arrays of `a` are updated in
 * place; a->vel / tri_idx / ns_cell / map_cell / flags are outputs (may be NULL).
 * Follows ECSystem::update restricted to Physics + Seek (src/ECS.cpp:65-113); see SURVEY Appendix A.
 */
int32_t rts_oracle_step(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                        const RtsWallTable* W, const RtsUnitType* types, uint32_t n_types, const RtsPathTable* paths,
                        uint32_t n, RtsAgentView* a, const uint8_t* ghost, float* fin_area) {
    acos_init();
    (void)n_types;
    ns_grid g;
    uint32_t* cell_of = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    build_grid(P, n, a->r, &g, cell_of);

    /* snapshot of what neighbours see: positions/states at frame start (components are copied into the
       cells before any agent is updated, Context.cpp:31-43) */
    float* r0 = (float*)malloc(sizeof(float) * 2 * (n ? n : 1));
    uint8_t* s0 = (uint8_t*)malloc(n ? n : 1);
    memcpy(r0, a->r, sizeof(float) * 2 * n);
    memcpy(s0, a->state, n);
    /* alignment extension (RtsParams.enable_alignment): neighbours see the merged velocity of the previous frame */
    float* v0 = NULL;
    if (P->enable_alignment && a->vel) {
        v0 = (float*)malloc(sizeof(float) * 2 * (n ? n : 1));
        memcpy(v0, a->vel, sizeof(float) * 2 * n);
    }

    /* ---- arrival (SeekSystem::finalizeSeek, SeekSystem.cpp:235-275) with frame-start snapshot semantics: see
       RtsParams.enable_arrival. stop[i] = agent i stops this frame. The cone rule (neighboursInFrontReachedEnd, :153-179)
       needs STANDING agents that are still members of the group; stopSeeking always removes an agent from its group
       (:136-144, CommandGroupManager.h:141-161), so it never fires and is not evaluated. */
    uint8_t* stop = (uint8_t*)calloc(n ? n : 1, 1);
    if (P->enable_arrival && fin_area && a->path_id) {
        const float area_unit = (float)3.14159265358979323846 * P->rhard * P->rhard; /* M_PIf * RHARD * RHARD */
        for (uint32_t i = 0; i < n; ++i) {
            const int ghost_i = ghost && ghost[i];
            if (ghost_i && !g_slab_arrival) continue;
            const int32_t pid = a->path_id[i];
            if (s0[i] != RTS_MOVING || pid < 0) continue;
            const uint32_t p0 = paths->offsets[pid], last = paths->offsets[pid + 1] - 1;
            const uint32_t wpi = a->waypoint ? (uint32_t)a->waypoint[i] : 0u;
            const uint32_t w0 = p0 + wpi < last ? p0 + wpi : last;
            v2 r_next = V(a->r_next[2 * i], a->r_next[2 * i + 1]);
            if (r_next.x != r_next.x) r_next = V(paths->waypoints[2 * w0], paths->waypoints[2 * w0 + 1]);
            const v2 path_end = V(paths->path_end[2 * pid], paths->path_end[2 * pid + 1]);
            if (!vequal(path_end, r_next)) continue;                          /* SeekSystem.cpp:106 */
            const int32_t g_i = paths->group ? paths->group[pid] : pid;
            const v2 r = V(r0[2 * i], r0[2 * i + 1]);
            const float finish_radius = sqrtf(fin_area[g_i] / (float)3.14159265358979323846);
            const float dist_to_end = sqrtf(vdist2(path_end, r));
            if (!(dist_to_end < 2 * a->radius[i] + finish_radius)) continue; /* :258 */
            if (!ghost_i) stop[i] = 1;
            const uint32_t ci = cell_of[i];
            const int cx = (int)(ci % (uint32_t)P->ns_nx), cy = (int)((ci / (uint32_t)P->ns_nx) % (uint32_t)P->ns_ny);
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy) {
                    if (cx + dx < 0 || cx + dx >= P->ns_nx || cy + dy < 0 || cy + dy >= P->ns_ny) continue;
                    const uint32_t cj = (uint32_t)(cx + dx) + (uint32_t)(cy + dy) * (uint32_t)P->ns_nx;
                    for (uint32_t k = g.cell_start[cj]; k < g.cell_start[cj + 1]; ++k) {
                        const uint32_t j = g.cell_items[k];
                        if (j == i || (ghost && ghost[j]) || s0[j] != RTS_MOVING) continue;
                        const int32_t pj = a->path_id[j];
                        if (pj < 0 || (paths->group ? paths->group[pj] : pj) != g_i) continue;
                        const v2 dr = vsub(V(r0[2 * j], r0[2 * j + 1]), r);
                        const float rr = a->radius[i] + a->radius[j];
                        if (!(vnorm2(dr) < 10.f * rr * rr)) continue;         /* seek neighbour list, Strategy.cpp:109-111 */
                        if (sqrtf(vnorm2(vsub(r, V(r0[2 * j], r0[2 * j + 1])))) < P->finish_radius) stop[j] = 1; /* :266-270 */
                    }
                }
        }
        for (uint32_t i = 0; i < n; ++i)
            if (stop[i]) {
                const int32_t pid = a->path_id[i];
                fin_area[paths->group ? paths->group[pid] : pid] += area_unit; /* CommandGroupManager.h:150 */
                if (g_stop_counts) g_stop_counts[paths->group ? paths->group[pid] : pid] += 1u;
            }
    }

    int32_t rc = RTS_OK;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
        const uint32_t i = (uint32_t)ii;
        if (ghost && ghost[i]) continue;
        const v2 r = V(r0[2 * i], r0[2 * i + 1]);
        const float radius = a->radius[i], mass_i = a->mass[i];
        const int state_i = s0[i];
        const RtsUnitType* ut = &types[a->unit_type ? a->unit_type[i] : 0];
        uint32_t flags = 0;

        /* ---- P1+P2: neighbour pass + applyForces (Strategy.cpp:10-68, PhysicsSystem.cpp:54-117) ---- */
        const uint32_t ci = cell_of[i];
        const int cx = (int)(ci % (uint32_t)P->ns_nx), cy = (int)((ci / (uint32_t)P->ns_nx) % (uint32_t)P->ns_ny);
        uint32_t cells[9];
        int ncells = 0;
        for (int dx = -1; dx <= 1; ++dx)        /* SearchGrid::calcNearestCells, Grid.cpp:103-118 */
            for (int dy = -1; dy <= 1; ++dy) {
                const uint32_t cj = ci + dx + dy * P->ns_nx;
                if (cx + dx >= 0 && cx + dx < P->ns_nx && cy + dy >= 0 && cy + dy < P->ns_ny && cj != ci) cells[ncells++] = cj;
            }
        cells[ncells++] = ci;                   /* centre last, Strategy.cpp:34 */

        v2 repulsion = V(0, 0), push = V(0, 0), scatter = V(0, 0), dr_nn = V(0, 0), align = V(0, 0);
        float n_neighbours = 0, n_align = 0;
        for (int c = 0; c < ncells; ++c) {
            for (uint32_t k = g.cell_start[cells[c]]; k < g.cell_start[cells[c] + 1]; ++k) {
                const uint32_t j = g.cell_items[k];
                const v2 dr = vsub(V(r0[2 * j], r0[2 * j + 1]), r);
                if (!(vnorm2(dr) < P->r_max2 && i != j)) continue;
                const float r_collision = a->radius[j] + radius;
                const float mass_j = a->mass[j];
                const int state_j = s0[j];
                const float r_collision_sq = r_collision * r_collision;
                const float d_surfs = vnorm2(dr);
                const float x = r_collision_sq / d_surfs;
                if (x > 0.8f) {
                    const float m3 = 3.f * x * x * x;
                    const float mag = (m3 < 50000.0f) ? m3 : 50000.0f; /* std::min(50000.0f, m3) */
                    repulsion = vadd(repulsion, vdiv(vmul(vmul(vdiv(vneg(dr), vnorm(dr)), mag), mass_j), (mass_i + mass_j)));
                }
                if (state_j == RTS_MOVING && state_i == RTS_STANDING && x > 0.75f)
                    push = vadd(push, vdiv(vmul(vmul(vneg(dr), x), mass_j), (mass_i + mass_j)));
                if (state_j == RTS_MOVING && x > r_collision_sq / (P->rscatter * P->rscatter)) {
                    dr_nn = vadd(dr_nn, dr);
                    n_neighbours++;
                }
                /* extension: the predicate of the commented-out block PhysicsSystem.cpp:87-91 (without the group test) */
                if (v0 && state_i == RTS_MOVING && state_j == RTS_MOVING && x > r_collision_sq / (P->ralign * P->ralign)) {
                    align = vadd(align, vsub(V(v0[2 * j], v0[2 * j + 1]), V(v0[2 * i], v0[2 * i + 1])));
                    n_align++;
                }
            }
        }
        dr_nn = vdiv(dr_nn, n_neighbours);
        if (n_neighbours > 1) scatter = vmul(vdiv(dr_nn, vnorm(dr_nn)), ut->max_speed);
        v2 inertia = V(a->inertia[2 * i], a->inertia[2 * i + 1]);
        inertia = vadd(inertia, vadd(vadd(vmul(repulsion, P->mul_repulse), vmul(push, P->mul_push)), vmul(scatter, P->mul_scatter)));
        if (v0 && n_align > 0) inertia = vadd(inertia, vmul(vdiv(align, n_align), P->mul_align));

        /* ---- P3: PhysicsSystem::update body, PhysicsSystem.cpp:22-44 (vel is 0 on entry) ---- */
        v2 vp = V(0, 0);
        float speed = vnorm(vp);
        if (speed > P->sys_max_speed) vp = vdiv(vp, speed * P->sys_max_speed);
        vp = vadd(vp, inertia);
        const float max_vel = ut->max_speed * P->mul_velocity;
        speed = vnorm(vp);
        if (speed > max_vel) vp = vmul(vp, max_vel / speed);
        const float lm = ut->lambda * P->mul_decay;
        const float decay = (lm < 1.f) ? lm : 1.f; /* std::min(1.f, lm) */
        inertia = vsub(inertia, vmul(inertia, decay));

        /* ---- P4: wall pass #1, PhysicsSystem.cpp:130-162 ---- */
        avoid_ctx ac = {&vp, &inertia};
        contact_walls(P, W, r, radius, avoid_one, &ac);

        /* ---- S1: SeekSystem::update per-agent body, SeekSystem.cpp:76-111 ---- */
        v2 vs = V(0, 0);
        float angle = a->angle[i], angle_vel = a->angle_vel[i];
        int state = state_i;
        int32_t wp = a->waypoint ? a->waypoint[i] : 0;
        v2 r_next = V(a->r_next[2 * i], a->r_next[2 * i + 1]);
        int32_t pid = a->path_id ? a->path_id[i] : -1;
        if (stop[i]) { /* stopSeeking, SeekSystem.cpp:136-144: no seek velocity this frame, out of the group */
            state = RTS_STANDING;
            angle_vel = 0;
            pid = -1;
            a->path_id[i] = -1;
        }
        if (state == RTS_MOVING && pid >= 0) {
            const uint32_t p0 = paths->offsets[pid], p1 = paths->offsets[pid + 1];
            const uint32_t last = p1 - 1;
            const uint32_t w0 = p0 + (uint32_t)wp < last ? p0 + (uint32_t)wp : last;
            const uint32_t w1 = w0 + 1 < last ? w0 + 1 : last;
            if (r_next.x != r_next.x) r_next = V(paths->waypoints[2 * w0], paths->waypoints[2 * w0 + 1]);
            const v2 r_nn = V(paths->waypoints[2 * w1], paths->waypoints[2 * w1 + 1]);
            const RtsEdge* portal_next = &paths->portals[w0];
            const v2 path_end = V(paths->path_end[2 * pid], paths->path_end[2 * pid + 1]);

            const v2 dr_to_target = vsub(r_next, r);
            /* turnTowards, SeekSystem.cpp:344-385 (the local `angle` edit is not communicated, SeekSystem.cpp:322-342) */
            {
                float ang = angle;
                if (ang > 180) ang -= 360;
                if (ang < -180) ang += 360;
                float desired;
                if (vnorm2(dr_to_target) == 0) desired = ut->desired_angle;
                else desired = table_angle(dr_to_target);
                float d_angle = desired - ang;
                if (d_angle > 180) d_angle = -360 + d_angle;
                if (d_angle < -180) d_angle = 360 + d_angle;
                angle_vel = (d_angle < ut->turn_rate) ? d_angle : ut->turn_rate; /* std::min({turn_rate, d_angle}) */
                if (d_angle < 0) angle_vel = (d_angle > -ut->turn_rate) ? d_angle : -ut->turn_rate;
                if (fabsf(d_angle) < ut->turn_rate) angle_vel = 0;
            }
            const float norm_dr = vnorm(dr_to_target);
            if (norm_dr > radius) vs = vmul(vdiv(dr_to_target, norm_dr), ut->seek_max_speed);

            /* needsUpdating, SeekSystem.cpp:181-233 */
            int upd;
            if (vequal(r_next, path_end)) upd = 0;
            else {
                const v2 t = vsub(r_nn, r_next);
                const int can_move = vdot(vneg(dr_to_target), t) > 0;
                upd = (norm_dr < P->rhard && can_move) || veq(r_next, r_nn);
                const v2 cp = closest_point_on_edge(r, portal_next);
                if (norm_dr < radius || vdot(dr_to_target, t) < 0) upd = 1;
                else {
                    if (vdist2(cp, r) < vdist2(r, r_next) && portal_next->l > 0 &&
                        vdist2(cp, r) < 16 * P->rhard * P->rhard && vnorm2(t) != 0) {
                        r_next = cp;
                        upd = 0;
                    } else if (vequal(r_next, r_nn) && !veq(r_next, path_end)) upd = 1;
                }
            }
            if (upd) { /* updateTarget, SeekSystem.cpp:277-304 */
                const v2 dn = vsub(r_nn, r);
                const float ndn = vnorm(dn);
                r_next = r_nn;
                if (p0 + (uint32_t)wp < last) wp++;
                if (ndn > 0) vs = vmul(vdiv(dn, ndn), ut->seek_max_speed);
                if (!veq(r_nn, path_end)) flags |= RTS_FLAG_REPLAN;
            }
            /* finalizeSeek (SeekSystem.cpp:235-275) is the order-dependent arrival cascade: next-row f2 */
        }

        /* ---- M: merge + wall pass #2 + integrate (PhysicsSystem.cpp:118-127, SeekSystem.cpp:322-342, ECS.cpp:92-113) ---- */
        v2 v = vadd(vadd(V(0, 0), vp), vs);
        avoid_ctx ac2 = {&v, &inertia};
        contact_walls(P, W, r, radius, avoid_one, &ac2);
        const v2 r_new = vadd(r, vmul(v, P->dt));
        angle = angle + angle_vel;

        a->r[2 * i] = r_new.x; a->r[2 * i + 1] = r_new.y;
        if (a->vel) { a->vel[2 * i] = v.x; a->vel[2 * i + 1] = v.y; }
        a->inertia[2 * i] = inertia.x; a->inertia[2 * i + 1] = inertia.y;
        a->angle[i] = angle; a->angle_vel[i] = angle_vel;
        a->state[i] = (uint8_t)state;
        if (a->waypoint) a->waypoint[i] = wp;
        a->r_next[2 * i] = r_next.x; a->r_next[2 * i + 1] = r_next.y;
        if (a->tri_idx && (P->walk_all || state == RTS_MOVING)) {
            a->tri_idx[i] = find_triangle(P, tris, n_tris, cell2tri, r_new.x, r_new.y, &flags);
            if (a->tri_idx[i] == 0xFFFFFFFFu) flags |= RTS_FLAG_WALK_FAILED;   /* no containing triangle at all */
            else flags &= ~(uint32_t)RTS_FLAG_WALK_FAILED;
        }
        if (a->ns_cell) a->ns_cell[i] = (uint32_t)coord_to_cell(r_new.x, r_new.y, P->ns_nx, P->ns_ny, P->ns_cell_size);
        if (a->map_cell) a->map_cell[i] = (uint32_t)coord_to_cell(r_new.x, r_new.y, P->map_nx, P->map_ny, P->map_cell_size);
        if (a->flags) a->flags[i] = flags;
    }
    free(v0); free(r0); free(s0); free(cell_of); free(g.cell_start); free(g.cell_items); free(stop);
    return rc;
}

/* NeighbourSearcherContext::getNeighboursIndsFull (NeighbourSearcherContext.h:44-92): indices of the agents within r_max of
   the point q, in the reference's visit order (cell rows j, columns i, ascending component index inside a cell); the cells
   searched are rows cell.y +- (ceil(r_max/cs)+1) and columns cell.x +- (j_max - j_min), clipped to the grid. Used by the game
   for click selection on the physics grid (Game.cpp:324). Returns the count; writes at most cap indices. */
uint32_t rts_oracle_query_radius(const RtsParams* P, uint32_t n, const float* r, float qx, float qy, float r_max, uint32_t* out, uint32_t cap) {
    ns_grid g;
    uint32_t* cell_of = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    build_grid(P, n, r, &g, cell_of);
    const float r_max2 = r_max * r_max;
    const uint32_t cell_ind = (uint32_t)coord_to_cell(qx, qy, P->ns_nx, P->ns_ny, P->ns_cell_size);
    const int ccx = (int)(cell_ind % (uint32_t)P->ns_nx), ccy = (int)(cell_ind / (uint32_t)P->ns_nx);   /* Grid::cellCoords(int), Grid.cpp:60-64 */
    int j_max = ccy + (int)ceilf(r_max / P->ns_cell_size) + 1;
    int j_min = ccy - (int)ceilf(r_max / P->ns_cell_size) - 1;
    if (j_max > P->ns_ny - 1) j_max = P->ns_ny - 1;
    if (j_min < 0) j_min = 0;
    const int delta_j_max = j_max - j_min;
    uint32_t cnt = 0;
    for (int j = j_min; j <= j_max; ++j) {
        int i_max = ccx + delta_j_max, i_min = ccx - delta_j_max;
        if (i_max > P->ns_nx - 1) i_max = P->ns_nx - 1;
        if (i_min < 0) i_min = 0;
        for (int i = i_min; i <= i_max; ++i) {
            const uint32_t c = (uint32_t)i + (uint32_t)j * (uint32_t)P->ns_nx;
            for (uint32_t k = g.cell_start[c]; k < g.cell_start[c + 1]; ++k) {
                const uint32_t a = g.cell_items[k];
                const float d2 = vdist2(V(r[2 * a], r[2 * a + 1]), V(qx, qy));
                if (d2 <= r_max2) { if (cnt < cap) out[cnt] = a; cnt++; }
            }
        }
    }
    free(cell_of); free(g.cell_start); free(g.cell_items);
    return cnt;
}

/* NeighbourSearcherStrategyAttack::execute (NeighbourSearcherStrategy.cpp:125-177) as written: over the cells of
   calcNearestCells + the own cell, in that order and ascending component index inside a cell, EVERY agent of another player
   with dist^2 < range2 (the reference never lowers min_enemy_dist_sq from its initial Geometry::BOX[0] = 2560) overwrites slot
   0 — so the "nearest enemy" it reports is the LAST such agent in visit order. out[i] = that index, or -1.
   Pinned: the verbatim strategy runs headless through NeighbourSearcherContext<AttackComponent, int> (ref_harness.cpp::
   refh_attack_ns; only AttackSystem.h itself needs glm, and the strategy's translation unit is built without it) and
   tests/test_oracle_vs_ref.py compares every agent's answer on random two- and three-player crowds. */
void rts_oracle_attack_targets(const RtsParams* P, uint32_t n, const float* r, const uint8_t* player, float range2, int32_t* out) {
    ns_grid g;
    uint32_t* cell_of = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    build_grid(P, n, r, &g, cell_of);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ci = cell_of[i];
        const int cx = (int)(ci % (uint32_t)P->ns_nx), cy = (int)((ci / (uint32_t)P->ns_nx) % (uint32_t)P->ns_ny);
        uint32_t cells[9];
        int ncells = 0;
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) {
                const uint32_t cj = ci + dx + dy * P->ns_nx;
                if (cx + dx >= 0 && cx + dx < P->ns_nx && cy + dy >= 0 && cy + dy < P->ns_ny && cj != ci) cells[ncells++] = cj;
            }
        cells[ncells++] = ci;
        int32_t target = -1;
        for (int c = 0; c < ncells; ++c)
            for (uint32_t k = g.cell_start[cells[c]]; k < g.cell_start[cells[c] + 1]; ++k) {
                const uint32_t j = g.cell_items[k];
                const float dist_sq = vdist2(V(r[2 * j], r[2 * j + 1]), V(r[2 * i], r[2 * i + 1]));
                if (dist_sq < range2 && player[i] != player[j]) target = (int32_t)j;
            }
        out[i] = target;
    }
    free(cell_of); free(g.cell_start); free(g.cell_items);
}

/* physics neighbour list of agent i in visit order (for the neighbour-set test T3) */
uint32_t rts_oracle_neighbours(const RtsParams* P, uint32_t n, const float* r, uint32_t i, uint32_t* out, uint32_t cap) {
    ns_grid g;
    uint32_t* cell_of = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    build_grid(P, n, r, &g, cell_of);
    const uint32_t ci = cell_of[i];
    const int cx = (int)(ci % (uint32_t)P->ns_nx), cy = (int)((ci / (uint32_t)P->ns_nx) % (uint32_t)P->ns_ny);
    uint32_t cells[9];
    int ncells = 0;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const uint32_t cj = ci + dx + dy * P->ns_nx;
            if (cx + dx >= 0 && cx + dx < P->ns_nx && cy + dy >= 0 && cy + dy < P->ns_ny && cj != ci) cells[ncells++] = cj;
        }
    cells[ncells++] = ci;
    uint32_t cnt = 0;
    for (int c = 0; c < ncells; ++c)
        for (uint32_t k = g.cell_start[cells[c]]; k < g.cell_start[cells[c] + 1]; ++k) {
            const uint32_t j = g.cell_items[k];
            const v2 dr = vsub(V(r[2 * j], r[2 * j + 1]), V(r[2 * i], r[2 * i + 1]));
            if (vnorm2(dr) < P->r_max2 && i != j) { if (cnt < cap) out[cnt] = j; cnt++; }
        }
    free(cell_of); free(g.cell_start); free(g.cell_items);
    return cnt;
}

void rts_oracle_default_params(RtsParams* p, int32_t map_nx, int32_t map_ny) {
    memset(p, 0, sizeof(*p));
    p->map_nx = map_nx; p->map_ny = map_ny; p->map_cell_size = 5.f;
    p->ns_cell_size = 60.f; /* RHARD*20, PhysicsSystem.cpp:7 */
    p->ns_nx = (int32_t)((float)(map_nx * 5) / p->ns_cell_size) + 1; /* Context.cpp:70-71 */
    p->ns_ny = (int32_t)((float)(map_ny * 5) / p->ns_cell_size) + 1;
    p->r_max2 = 30.f;
    p->tri_nx = map_nx / 8; p->tri_ny = map_ny / 8; p->tri_cell_size = 40.f;
    p->mul_push = 0.1f; p->mul_repulse = 1.f; p->mul_scatter = 0.1f; p->mul_align = 1.f; p->mul_seek = 1.f;
    p->mul_decay = 1.f; p->mul_velocity = 2.f;
    p->dt = (float)(1. / 60.f); /* constexpr float dt = 1. / 60.f, ECS.h:18 */
    p->rhard = 3.f; p->rscatter = 50.f; p->sys_max_speed = 25.f; p->ralign = 50.f;
    p->finish_angle = 36; p->finish_radius = 50; p->finish_n_steps = 1; p->finish_cone_len = 10;
    p->enable_arrival = 0; p->walk_all = 1;
    p->slab_row_lo = 0; p->slab_row_hi = p->ns_ny;
}

int rts_oracle_threads(void) { return omp_get_max_threads(); }
void rts_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
/* Slab model of the arrival rule (tests/test_slabs_gloo.py): with `on`, agents marked as ghosts also EVALUATE the rule — they
   stop the owned agents among their same-group neighbours, never themselves (their owner does) — which is what every rank of
   the slab decomposition does for the halo copies it holds. `counts` (may be NULL): per-group number of agents stopped by the
   following steps, accumulated. */
void rts_oracle_set_slab_arrival(int on, uint32_t* counts) { g_slab_arrival = on; g_stop_counts = counts; }

void rts_oracle_walk_hops(const RtsParams* P, const RtsTriangle* tris, uint32_t n_tris, const uint32_t* cell2tri,
                          const float* xy, uint32_t n, uint32_t* hops_out) {
    for (uint32_t i = 0; i < n; ++i) {
        g_last_hops = 0;
        (void)find_triangle(P, tris, n_tris, cell2tri, xy[2 * i], xy[2 * i + 1], 0);
        hops_out[i] = g_last_hops;
    }
}
