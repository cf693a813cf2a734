// Device-side arithmetic of the agent update. Every function restates one reference function with
// the same operation order so that, compiled with --fmad=false (and nvcc's default IEEE division and
// square root), the results are bit-identical to an IEEE build of the reference on the CPU.
// Paths cited are relative to the reference root.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rts_gpu.h"

namespace rts {

// ---- constant tables of one handle (device pointers + scalars), passed to kernels by value ----
struct MapDev {
    // CDT: 3 x int4 per triangle {v0x v0y v1x v1y | v2x v2y n0 n1 | n2 flags 0 0}
    const int4* __restrict__ tris;
    uint32_t n_tris;
    const uint32_t* __restrict__ cell2tri;
    // walls: per line {first, end} into one edge pool, the lines of the four orientations concatenated (orientation o starts
    // at line_base[o]). Explicit ranges instead of CSR offsets: rts_update_map_region appends a changed line at the end of
    // the pool and redirects the line, nothing else moves.
    const uint2* __restrict__ line_se;
    const float4* __restrict__ edge_ft;     // {from.x from.y t.x t.y}
    const float* __restrict__ edge_l;
    uint32_t line_base[4];
    uint32_t n_lines[4];
    // paths: 2 x float4 per waypoint {w.x w.y p.from.x p.from.y | p.t.x p.t.y p.l 0}
    const uint32_t* __restrict__ path_off;
    const float4* __restrict__ waypoints;
    const float2* __restrict__ path_end;
    uint32_t n_paths;
    const int32_t* __restrict__ path_group;   // command group of each path
    uint32_t n_groups;
    float* area_read;    // CommandGroupManager2::group2finished_surface_area_, value at the start of the frame
    float* area_acc;     // ... accumulating this frame's stops
    // slab handles: the finished area is the same float sum, rebuilt on every rank from COUNTS — stop_local = stops this rank has
    // claimed (cumulative), stop_global = their sum over the ranks after the last exchange, stop_applied = how many of those
    // k_area_sync has already added (one by one, pi*RHARD^2 each: the float sequence of the single-domain atomicAdds)
    uint32_t* stop_local;
    uint32_t* stop_global;
    uint32_t* stop_applied;
    const RtsUnitType* __restrict__ unit_types;
    uint32_t n_types;
    const float* __restrict__ acos_table;   // AcosTable<1000>, built on the host (src/core.h:115-124)
    // conservative wall pre-filter: one bit per block of WALL_BLOCK map units, set if any wall lies within
    // wall_filter_radius (+1) of the block; agents with a larger radius always run the full query
    const uint32_t* __restrict__ wall_bits;
    int32_t wall_bx, wall_by;
    float wall_filter_radius;
    float inv_wall_block;     // 1 / block size of the bitmap (one MapGrid cell: 5 units)
    int32_t coords_fit_i32;   // 1: every triangle-walk product fits int32 (the reference's own arithmetic)
    // acceleration of findTriangle's linear-scan fallback (lowest triangle index containing q): per cache cell the
    // ascending list of "small" triangles whose bounding box touches the cell, plus one ascending list of the few
    // "large" triangles (bounding box over SCAN_LARGE_CELLS cells)
    const uint2* __restrict__ scan_se;      // per cache cell {first, end} into scan_list (same pool scheme as the walls)
    const uint32_t* __restrict__ scan_list;
    const uint32_t* __restrict__ scan_large;
    uint32_t n_scan_large;
    const uint32_t* __restrict__ scan_outer;   // triangles reaching beyond the cache grid (the super-triangle fan)
    uint32_t n_scan_outer;
    // findTriangle's answers for the integer points ON the far border lines of the cache grid (x == W or y == H): the cache cell
    // of such a point wraps to the opposite side of the map (Grid.cpp:18-24) and the reference's walk crosses the whole map —
    // hundreds of dependent hops for every agent that is just crossing the line (config 5: ~0.28 ms per frame for a handful of
    // agents). Filled after every map change by that very walk (k_border_cache), so the answers are the reference's own.
    const uint32_t* __restrict__ border_top;     // [border_gw + 1]: q = (x, border_gh)
    const uint32_t* __restrict__ border_right;   // [border_gh + 1]: q = (border_gw, y)
    int32_t border_gw, border_gh;                // extent of the cache grid in map units (integers); pointers null = no cache
    // per path {offset, end offset, path_end.x, path_end.y} in one 16-byte word (same numbers as path_off / path_end)
    const float4* __restrict__ path_hdr;
    // RN(1 / ns_cell_size), RN(1 / map_cell_size): used by floor_div() to skip the IEEE division on its common path
    float inv_ns_cs, inv_map_cs;
};
constexpr uint32_t SCAN_LARGE_CELLS = 1024;

struct v2 {
    float x, y;
};
__device__ __forceinline__ v2 V(float x, float y) { return v2{x, y}; }
// sf::Vector2 operators are component-wise (src/Utils/Vector.inl)
__device__ __forceinline__ v2 operator+(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ v2 operator-(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ v2 operator-(v2 a) { return V(-a.x, -a.y); }
__device__ __forceinline__ v2 operator*(v2 a, float s) { return V(a.x * s, a.y * s); }
__device__ __forceinline__ v2 operator/(v2 a, float s) { return V(a.x / s, a.y / s); }
// src/core.h:99-113
__device__ __forceinline__ float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float norm2(v2 a) { return dot(a, a); }
__device__ __forceinline__ float norm(v2 a) { return sqrtf(norm2(a)); }
__device__ __forceinline__ float dist2(v2 a, v2 b) { return dot(a - b, a - b); }
__device__ __forceinline__ bool vequal(v2 a, v2 b) { return dist2(a, b) < 0.001f; }
__device__ __forceinline__ bool veq(v2 a, v2 b) { return a.x == b.x && a.y == b.y; }

// ---- Grid::coordToCell, src/Utils/Grid.cpp:18-24 ------------------------------------------------
// floorf(x / cs), bit for bit, without the IEEE division on the common path. q~ = RN(x * RN(1/cs)) and the reference's
// RN(x / cs) both lie within 1.8e-7 |q| of the real quotient, so whenever q~ is further than 2.5e-7 |q~| from the nearest
// integers no integer separates the two and their floors agree; otherwise (and for NaN / |q| >= 2^23, where the
// fractional part is 0) the division is done. inv_cs must be RN(1/cs).
__device__ __forceinline__ float floor_div(float x, float cs, float inv_cs) {
    const float q = x * inv_cs;
    float f = floorf(q);
    const float d = q - f;
    const float m = fabsf(q) * 2.5e-7f;
    if (!(d > m && d < 1.f - m)) f = floorf(x / cs);
    return f;
}
// size_t(floor(x/cs)) % n.  The common case (0 <= floor < 2^31) stays in 32 bits; anything else takes
// the 64-bit path that reproduces x86-64's float -> size_t conversion (signed convert, wrap mod 2^64).
__device__ __forceinline__ uint32_t cell_of_floor(float f, int32_t n) {
    if (f >= 0.f && f < 2147483648.f) {
        const uint32_t u = (uint32_t)f;
        return u < (uint32_t)n ? u : u % (uint32_t)n;   // (inside the map nothing wraps: no division)
    }
    const uint64_t u = (uint64_t)(int64_t)f;   // cvttss2si semantics
    return (uint32_t)(u % (uint64_t)n);
}
__device__ __forceinline__ uint32_t axis_cell(float x, float cs, int32_t n) { return cell_of_floor(floorf(x / cs), n); }
__device__ __forceinline__ uint32_t axis_cell(float x, float cs, float inv_cs, int32_t n) { return cell_of_floor(floor_div(x, cs, inv_cs), n); }
__device__ __forceinline__ uint32_t coord_to_cell(float x, float y, int32_t nx, int32_t ny, float cs) {
    return axis_cell(x, cs, nx) + axis_cell(y, cs, ny) * (uint32_t)nx;
}

// ---- approximations of the tolerance mode (RtsParams.precision = RTS_PRECISION_FAST) only ---------------------------
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// ... with one Newton step (relative error ~1e-7): the once-per-agent normalisations, whose result is scaled by a speed
__device__ __forceinline__ float rsqrt_nr(float x) {
    const float y = rsqrt_fast(x);
    return y * __fmaf_rn(-0.5f * x * y, y, 1.5f);
}

// ---- AcosTable<1000>::angle, src/core.h:139-150 ---------------------------------------------------
__device__ __forceinline__ float table_angle(const float* __restrict__ table, v2 v) {
    const float dot_value = (v.x * 1.f + v.y * 0.f) / norm(v) + 1.0f;
    int bin = (int)floorf(dot_value / 2.f * 1000);
    if (bin >= 1000) bin = 999;
    if (bin < 0) bin = 0;
    return __ldg(table + bin) * (2.f * (v.y > 0) - 1.f);
}

// ---- Triangulation::findTriangle, src/PathFinding/Triangulation.cpp:1497-1567 ----------------------
struct iv2 {
    int32_t x, y;
};
// sign(), Triangulation.h:73-76: integer arithmetic, then ->float.  WIDE = false is the reference's own int32
// arithmetic (used when the map is small enough that nothing overflows); WIDE = true forms the products in 64 bits,
// identical wherever int32 does not overflow (on 8192² maps the reference overflows, which is UB there).
template <bool WIDE>
__device__ __forceinline__ float tri_sign(iv2 p1, iv2 p2, iv2 p3) {
    if (WIDE) return (float)((int64_t)(p1.x - p3.x) * (p2.y - p3.y) - (int64_t)(p2.x - p3.x) * (p1.y - p3.y));
    return (float)((p1.x - p3.x) * (p2.y - p3.y) - (p2.x - p3.x) * (p1.y - p3.y));
}
template <bool WIDE>
__device__ __forceinline__ bool in_triangle(iv2 q, iv2 a, iv2 b, iv2 c) {  // isInTriangle, Triangulation.h:83-96
    const float d1 = tri_sign<WIDE>(q, a, b), d2 = tri_sign<WIDE>(q, b, c), d3 = tri_sign<WIDE>(q, c, a);
    const bool has_neg = (d1 < 0) || (d2 < 0) || (d3 < 0);
    const bool has_pos = (d1 > 0) || (d2 > 0) || (d3 > 0);
    return !(has_neg && has_pos);
}
template <bool WIDE>
__device__ __forceinline__ float idot(iv2 a, iv2 b) {  // dot<Vertex>, core.h:99
    if (WIDE) return (float)((int64_t)a.x * b.x + (int64_t)a.y * b.y);
    return (float)(a.x * b.x + a.y * b.y);
}
// RN(N / D) in [0, 1] for finite floats N, D — decided without dividing, exactly:
//   D == 0          -> the quotient is +-inf or NaN: false
//   N == 0          -> +-0: true
//   opposite signs  -> negative (no underflow: |N| >= 1 here... any nonzero finite quotient keeps its sign): false
//   same signs      -> RN(q) <= 1  <=>  q <= 1 + 2^-24 (the tie rounds to even = 1). |N| <= |D| gives q <= 1. |N| > |D|
//                      means |N| >= |D| + ulp(|D|), and ulp(|D|)/|D| > 2^-24 strictly (|D| < 2^(e+1)), so q > 1 + 2^-24.
//   Hence the quotient is in [0,1]  <=>  D != 0 and (N == 0 or (sign N == sign D and |N| <= |D|)).
__device__ __forceinline__ bool quotient_in_unit_interval(float N, float D) {
    return (D > 0.f && N >= 0.f && N <= D) || (D < 0.f && N <= 0.f && N >= D);
}
template <bool WIDE>
__device__ __forceinline__ bool lines_intersect(iv2 v11, iv2 v12, iv2 v21, iv2 v22) {  // Triangulation.cpp:1805-1818
    const iv2 dr{v21.x - v11.x, v21.y - v11.y};
    const iv2 t1{v12.x - v11.x, v12.y - v11.y};
    const iv2 t2{v22.x - v21.x, v22.y - v21.y};
    const iv2 n1{t1.y, -t1.x};
    const iv2 n2{t2.y, -t2.x};
    // alpha1 = +dot(dr,n2)/dot(t1,n2), alpha2 = -dot(dr,n1)/dot(t2,n1), both required in [0,1]; the dots are integers
    // converted to float as in the reference (dot<Vertex> returns float, core.h:99), the two IEEE divisions are replaced
    // by the equivalent comparisons above
    return quotient_in_unit_interval(+idot<WIDE>(dr, n2), idot<WIDE>(t1, n2)) &&
           quotient_in_unit_interval(-idot<WIDE>(dr, n1), idot<WIDE>(t2, n1));
}

__device__ __forceinline__ void load_tri_verts(const MapDev& M, uint32_t t, iv2& a, iv2& b, iv2& c, int4& w1) {
    const int4 w0 = __ldg(M.tris + 3 * (size_t)t);
    w1 = __ldg(M.tris + 3 * (size_t)t + 1);
    a = iv2{w0.x, w0.y};
    b = iv2{w0.z, w0.w};
    c = iv2{w1.x, w1.y};
}

constexpr uint32_t TRI_NONE = 0xFFFFFFFFu;
constexpr uint32_t TRI_NEED_SCAN = 0xFFFFFFFEu;   // the reference falls back to its linear scan here

// Walk part of findTriangle. Returns the triangle, TRI_NEED_SCAN where the reference would scan all triangles
// (Triangulation.cpp:1503-1510, 1539-1546; done warp-cooperatively by the caller), or TRI_NONE + flag on failure.
template <bool WIDE>
__device__ __forceinline__ uint32_t find_triangle_from(const MapDev& M, iv2 q, uint32_t t, uint32_t& flags);
template <bool WIDE>
__device__ __forceinline__ uint32_t find_triangle_walk(const MapDev& M, const RtsParams& P, iv2 q, uint32_t& flags) {
    const uint32_t cell = coord_to_cell((float)q.x, (float)q.y, P.tri_nx, P.tri_ny, P.tri_cell_size);
    return find_triangle_from<WIDE>(M, q, __ldg(M.cell2tri + cell), flags);
}
// ... from an explicit start triangle: the cache entry of q's cell (from_last_found == false, :1500-1511), or the
// previously found triangle (from_last_found == true, used by updateCellGrid)
template <bool WIDE>
__device__ __forceinline__ uint32_t find_triangle_from(const MapDev& M, iv2 q, uint32_t t, uint32_t& flags) {
    if (t == TRI_NONE) return TRI_NEED_SCAN;
    iv2 a, b, c;
    int4 w1;
    load_tri_verts(M, t, a, b, c, w1);
    if (in_triangle<WIDE>(q, a, b, c)) return t;
    iv2 vc;
    {
        const iv2 g{(a.x + b.x + c.x) / 3, (a.y + b.y + c.y) / 3};  // integer division, :1524
        if (lines_intersect<WIDE>(g, q, a, b)) { t = (uint32_t)w1.z; vc = a; }
        else if (lines_intersect<WIDE>(g, q, b, c)) { t = (uint32_t)w1.w; vc = b; }
        else if (lines_intersect<WIDE>(g, q, c, a)) { t = (uint32_t)__ldg(M.tris + 3 * (size_t)t + 2).x; vc = c; }
        else t = TRI_NONE;
    }
    if (t == TRI_NONE) return TRI_NEED_SCAN;
    const uint32_t hop_limit = 4 * M.n_tris + 16;
    for (uint32_t hops = 0;; ++hops) {  // fan walk around vc, :1548-1563
        const int4 w2 = __ldg(M.tris + 3 * (size_t)t + 2);   // issued with the two vertex words: one round trip per hop
        load_tri_verts(M, t, a, b, c, w1);
        if (in_triangle<WIDE>(q, a, b, c)) return t;
        int k = -1;
        if (a.x == vc.x && a.y == vc.y) k = 0;
        else if (b.x == vc.x && b.y == vc.y) k = 1;
        else if (c.x == vc.x && c.y == vc.y) k = 2;
        if (k < 0 || hops >= hop_limit) {  // the reference would index out of bounds / never return: report instead
            flags |= RTS_FLAG_WALK_FAILED;
            return TRI_NONE;
        }
        const uint32_t n2 = (uint32_t)w2.x;
        const uint32_t nb0 = (uint32_t)w1.z, nb1 = (uint32_t)w1.w;
        const int kn = (k == 2) ? 0 : k + 1, kp = (k == 0) ? 2 : k - 1;
        const iv2 vn = (kn == 0) ? a : (kn == 1) ? b : c;
        const iv2 vp = (kp == 0) ? a : (kp == 1) ? b : c;
        if (lines_intersect<WIDE>(vc, q, vn, vp)) { t = (kn == 0) ? nb0 : (kn == 1) ? nb1 : n2; vc = vn; }
        else t = (k == 0) ? nb0 : (k == 1) ? nb1 : n2;
        if (t == TRI_NONE) { flags |= RTS_FLAG_WALK_FAILED; return t; }
    }
}

// findTriangle for the lanes of one warp. `want` says whether this lane has a query; every lane of `mask` must call.
// The rare linear scans are served by the whole warp: 32 triangles per iteration, lowest index wins.
template <bool WIDE, bool BORDER = true>
__device__ __forceinline__ uint32_t find_triangle_warp(const MapDev& M, const RtsParams& P, bool want, float fx, float fy,
                                                       uint32_t& flags, unsigned mask) {
    const iv2 q{(int32_t)fx, (int32_t)fy};  // static_cast<Vertex>(Vector2f): truncation, Triangulation.cpp:1588-1590
    uint32_t t = TRI_NONE;
    bool done = false;
    if (want && M.scan_se) {
        // A query outside the cache grid (an agent pushed over the map edge: the reference has no outer wall) gets its
        // start triangle from a cell index wrapped modulo the grid (Grid.cpp:18-24), i.e. from the far side of the map,
        // and the walk needs 100+ hops. Out there only the few triangles of the super-triangle fan exist. If the
        // query lies STRICTLY inside one of them, that triangle is the only one containing it, so every walk ends there.
        const float cxf = floorf((float)q.x / P.tri_cell_size), cyf = floorf((float)q.y / P.tri_cell_size);
        if (!(cxf >= 0.f && cyf >= 0.f && cxf < (float)P.tri_nx && cyf < (float)P.tri_ny)) {
            for (uint32_t k = 0; k < M.n_scan_outer && !done; ++k) {
                const uint32_t i = __ldg(M.scan_outer + k);
                iv2 a, b, c;
                int4 w1;
                load_tri_verts(M, i, a, b, c, w1);
                const float d1 = tri_sign<WIDE>(q, a, b), d2 = tri_sign<WIDE>(q, b, c), d3 = tri_sign<WIDE>(q, c, a);
                if ((d1 > 0 && d2 > 0 && d3 > 0) || (d1 < 0 && d2 < 0 && d3 < 0)) { t = i; done = true; }
            }
            if (BORDER && !done && M.border_top) {   // ON the far border line: the walk's own answer, recorded at the last map change
                if (q.y == M.border_gh && q.x >= 0 && q.x <= M.border_gw) { t = __ldg(M.border_top + q.x); done = true; }
                else if (q.x == M.border_gw && q.y >= 0 && q.y <= M.border_gh) { t = __ldg(M.border_right + q.y); done = true; }
            }
        }
    }
    if (want && !done) t = find_triangle_walk<WIDE>(M, P, q, flags);
    if (want && t == TRI_NEED_SCAN && M.scan_se) {
        // the reference scans every triangle in index order; a triangle that contains q has q inside its bounding box,
        // so only the triangles listed for q's cache cell (and the large ones) can be the answer
        const float cxf = floorf((float)q.x / P.tri_cell_size), cyf = floorf((float)q.y / P.tri_cell_size);
        if (cxf >= 0.f && cyf >= 0.f && cxf < (float)P.tri_nx && cyf < (float)P.tri_ny) {
            const uint32_t cell = (uint32_t)cxf + (uint32_t)cyf * (uint32_t)P.tri_nx;
            uint32_t best = TRI_NONE;
            const uint2 se = __ldg(M.scan_se + cell);
            for (uint32_t k = se.x, e = se.y; k < e; ++k) {
                const uint32_t i = __ldg(M.scan_list + k);
                iv2 a, b, c;
                int4 w1;
                load_tri_verts(M, i, a, b, c, w1);
                if (in_triangle<WIDE>(q, a, b, c)) { best = i; break; }
            }
            for (uint32_t k = 0; k < M.n_scan_large; ++k) {
                const uint32_t i = __ldg(M.scan_large + k);
                if (i >= best) break;
                iv2 a, b, c;
                int4 w1;
                load_tri_verts(M, i, a, b, c, w1);
                if (in_triangle<WIDE>(q, a, b, c)) { best = i; break; }
            }
            t = best;
        }
    }
    unsigned need = __ballot_sync(mask, want && t == TRI_NEED_SCAN);   // outside the cache grid: scan everything
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t n_lanes = (uint32_t)__popc(mask);                       // lanes that take part (exited lanes cannot)
    const uint32_t my = (uint32_t)__popc(mask & ((1u << lane) - 1u));      // rank of this lane among them
    while (need) {
        const int src = __ffs(need) - 1;
        const iv2 qq{__shfl_sync(mask, q.x, src), __shfl_sync(mask, q.y, src)};
        uint32_t found = TRI_NONE;
        for (uint32_t base = 0; base < M.n_tris; base += n_lanes) {
            const uint32_t i = base + my;
            bool hit = false;
            if (i < M.n_tris) {
                iv2 a, b, c;
                int4 w1;
                load_tri_verts(M, i, a, b, c, w1);
                hit = in_triangle<WIDE>(qq, a, b, c);
            }
            const unsigned hits = __ballot_sync(mask, hit);
            if (hits) {   // lowest triangle index = lowest participating lane that hit
                const int first = __ffs(hits) - 1;
                found = base + (uint32_t)__popc(mask & ((1u << first) - 1u));
                break;
            }
        }
        if ((int)lane == src) t = found;
        need &= need - 1;
    }
    return t;
}

// ---- Edges::calcContactEdgesO + isTouchingEdge, src/Edges.cpp:245-296, 128-143 ----------------------
__device__ __forceinline__ bool touching(float4 ft, float l, v2 r, float radius) {
    const v2 dr = V(ft.x, ft.y) - r;
    const v2 t = V(ft.z, ft.w);
    const float norm_dr_sq = dot(dr, dr);
    const float dot_dr_t = dot(dr, t);
    const float disc = dot_dr_t * dot_dr_t - norm_dr_sq + radius * radius;
    if (disc >= 0) {
        const float s = sqrtf(disc);
        const float a1 = -dot_dr_t + s, a2 = -dot_dr_t - s;
        return (a2 <= l && a2 > 0) || (a1 <= l && a1 > 0);
    }
    return false;
}

// Calls fn(t_x, t_y, edge_index) for every contact wall in the reference's order: orientation 0..3, line -D..+D,
// edge first..last.  Returns the number of contacts.
template <class Fn>
__device__ __forceinline__ int contact_walls(const MapDev& M, const RtsParams& P, v2 r, float radius, Fn fn) {
    const int ix0 = (int)(r.x / P.map_cell_size);  // Grid::cellCoords truncates, Grid.cpp:67-71
    const int iy0 = (int)(r.y / P.map_cell_size);
    const int D = (int)ceilf(radius / P.map_cell_size);
    int count = 0;
#pragma unroll 1
    for (int o = 0; o < 4; ++o) {
        const int base = (o == 0) ? iy0 : (o == 1) ? ix0 - iy0 + P.map_ny - 1 : (o == 2) ? ix0 : ix0 + iy0;
        const v2 dir = (o == 0) ? V(1.f, 0.f) : (o == 1) ? V(1.f, 1.f) : (o == 2) ? V(0.f, 1.f) : V(1.f, -1.f);
        const int lo = max(base - D, 0), hi = min(base + D, (int)M.n_lines[o] - 1);
        if (lo > hi) continue;
        const uint2* se = M.line_se + M.line_base[o];
#pragma unroll 1
        for (int L = lo; L <= hi; ++L) {
            const uint2 range = __ldg(se + L);
            const uint32_t beg = range.x, end = range.y;
            if (end != beg) {
                const float4* E = M.edge_ft + beg;
                const float* El = M.edge_l + beg;
                const int size = (int)(end - beg);
                // std::lower_bound (libstdc++ bisection) with comp(e, r) = dot(r - e.from, dir) > 0
                int first_i = 0, len = size;
                while (len > 0) {
                    const int half = len >> 1;
                    const int mid = first_i + half;
                    const float4 e = __ldg(E + mid);
                    if (dot(r - V(e.x, e.y), dir) > 0) { first_i = mid + 1; len = len - half - 1; }
                    else len = half;
                }
                int first = first_i - 1, last = first + 1;
                if (first == -1) { first = 0; last = 0; }
                if (first == size - 1) last = size - 1;
                for (int k = first; k <= last; ++k) {
                    const float4 e = __ldg(E + k);
                    if (touching(e, __ldg(El + k), r, radius)) { fn(e.z, e.w, beg + (uint32_t)k); ++count; }
                }
            }
        }
    }
    return count;
}

// Conservative pre-filter of the wall query: false only if no wall can touch a circle of `radius` at r
// (a contact needs dist(r, segment) <= radius; the bitmap marks every block within wall_filter_radius+1 of a wall).
__device__ __forceinline__ bool may_touch_wall(const MapDev& M, v2 r, float radius) {
    if (!M.wall_bits || !(radius <= M.wall_filter_radius)) return true;
    const float fx = r.x * M.inv_wall_block, fy = r.y * M.inv_wall_block;
    if (!(fx >= 0.f && fy >= 0.f && fx < (float)M.wall_bx && fy < (float)M.wall_by)) return true;   // outside the map (or NaN)
    const uint32_t b = (uint32_t)(int)fy * (uint32_t)M.wall_bx + (uint32_t)(int)fx;
    return (__ldg(M.wall_bits + (b >> 5)) >> (b & 31u)) & 1u;
}

// PhysicsSystem::avoidWall body for one wall, src/Systems/PhysicsSystem.cpp:146-158
__device__ __forceinline__ void avoid_one(float tx, float ty, v2& vel, v2& inertia) {
    const v2 n = V(ty, -tx);
    v2 away = n * dot(vel, n);
    if (dot(vel, n) < 0) vel = vel - away * 1.0f;
    away = n * dot(inertia, n);
    if (dot(inertia, n) < 0) inertia = inertia - away;
}

// closestPointOnEdge, src/core.h:282-296
__device__ __forceinline__ v2 closest_point_on_edge(v2 r, v2 from, v2 t, float l) {
    const v2 dr = r - from;
    const float pos = dot(dr, t);
    if (pos < l && pos > 0) return from + t * pos;
    else if (pos > l) return from + t * l;  // Edgef::to()
    return from;
}

}  // namespace rts
