// Surface semantics of mesh geometries: everything the GJK kernels leave open.
//
// The reference's meshes are SURFACES (Pinocchio loads a URDF <mesh> as a coal BVHModel,
// /root/reference/src/pyroboplan/models/panda.py:27): two meshes collide when two triangles come within
// the margin, a solid primitive collides with a mesh when it comes within the margin of a triangle, a
// shape wholly inside a mesh without reaching its surface does not collide with it.  The hot kernels
// bracket every mesh M between its hull H and an inner core K (pyroboplan_b200/model/surface.py):
//
//   d(H_a, H_b) <= d(M_a, M_b) <= d(K_a, K_b)        (when neither shape encloses the other)
//
// Boolean queries run GJK on the cores: "within the margin" is then certain, "farther than margin +
// eps_a + eps_b" (eps: how far a hull may lie outside its core) is certain, and the few items that
// converge in between are parked as UNDECIDED.  A core hit of a pair where one shape could lie inside the
// other's surface is parked as POSSIBLY ENCLOSED.  Distance queries run GJK on the hulls and park every
// item that could define the state's minimum.  The functions here decide the parked items:
//
//   lane level   converged GJK on hulls and cores (most undecided items have d_H == d_K, or d_H beyond the
//                margin); the centre-line and largest-face tests of the enclosure
//   warp level   the walk over all faces of the enclosing core, and the TRIANGLE STAGE: the triangles of
//                both meshes that can matter (slab along the separating direction, bounding spheres),
//                every pair of them through one converged GJK, the 32 lanes sharing the pairs.
#pragma once
#include <limits.h>

#include "pbp_gjk.cuh"

namespace pbp {

// Developer counters (build with -DPBP_DEBUG_STATS, read with pbp_debug_stats): off in the product build.
#if defined(PBP_DEBUG_STATS) && defined(__CUDACC__)
__device__ unsigned long long g_dbg[48];
#define DBG_ADD(i, v) atomicAdd(&g_dbg[i], (unsigned long long)(v))
#define DBG_MAX(i, v) atomicMax(&g_dbg[i], (unsigned long long)(v))
#define DBG_CLOCK() clock64()
#else
#define DBG_ADD(i, v) ((void)sizeof(i), (void)sizeof(v))
#define DBG_MAX(i, v) ((void)sizeof(i), (void)sizeof(v))
#define DBG_CLOCK() 0ll
#endif

// MODE_BOOL      verdict only; a state stops at its first certain hit, decided groups are skipped
// MODE_BOOL_ALL  verdict + first_pair / first_bad: every pair of every state is resolved
// MODE_DIST      verdict + min distance: pairs are resolved unless certainly farther than the bound
enum { MODE_BOOL = 0, MODE_BOOL_ALL = 1, MODE_DIST = 2 };

// lanes that share a warp-level loop: 32 on the device; the host build of this header (tests/host_emu)
// walks every loop with one lane and supplies one-lane versions of the warp intrinsics
#if defined(__CUDACC__)
constexpr int PBP_LANES = 32;
#else
constexpr int PBP_LANES = 1;
#endif

// bits set in Bins::group by the narrowphase on an item the refine kernel has to decide
constexpr int32_t SR_PARKED = 0x40000000;   // core hit of a pair with an enclosure flag
constexpr int32_t SR_DOUBT = 0x20000000;    // between the bounds (boolean) / may define the minimum (distance)
constexpr int32_t SR_BITS = SR_PARKED | SR_DOUBT;
// set by the item setup on an item of a pair with an enclosure flag whose shapes poke out of each other
// along the centre line: a core hit of this item is a hit, nothing to park (boolean queries)
constexpr int32_t SR_NOENC = 0x10000000;

#ifndef PBP_PAIR_FLAGS_DEFINED
#define PBP_PAIR_FLAGS_DEFINED
// per-pair flags (DevModel::pair_enclose, internal pair order)
enum : uint8_t {
    PAIR_ENCLOSE_B_IN_A = 1,   // `second` could lie wholly inside `first`, and `first` is a surface
    PAIR_ENCLOSE_A_IN_B = 2,   // `first` could lie wholly inside `second`, and `second` is a surface
    PAIR_ENCLOSE = 3,
    PAIR_SURFACE = 4,          // a geometry of the pair carries triangles
    PAIR_NOCORE = 8,           // ... and has no inner core: a hull hit proves nothing, the triangles decide every one
};

#endif

// how far the hulls of a pair may lie outside its cores, together
__device__ __forceinline__ float pair_eps(const DevGeom &GA, const DevGeom &GB) { return GA.core_eps + GB.core_eps; }

// where the calling kernel keeps the tables (shared or global memory)
struct SurfTables {
    const float4 *hverts; const uint16_t *hcells, *hlists;   // hulls
    const float4 *kverts; const uint16_t *kcells, *klists;   // inner cores
    const float4 *planes;
    const float4 *mverts;
    const int4 *tris;
    const float4 *tverts;   // [ntris][3] packed triangle vertices
};

__device__ __forceinline__ ShapeRef hull_shape(const SurfTables &t, const DevGeom &G) {
    ShapeRef S;
    S.shape = G.shape; S.p0 = G.par[0]; S.p1 = G.par[1]; S.p2 = G.par[2];
    S.verts = t.hverts + G.vert_off; S.nverts = G.vert_cnt;
    S.cell_off = G.map_off >= 0 ? t.hcells + G.map_off : nullptr; S.cell_list = t.hlists + G.list_off;
    return S;
}
__device__ __forceinline__ ShapeRef core_shape(const SurfTables &t, const DevGeom &G) {
    ShapeRef S;
    S.shape = G.shape; S.p0 = G.par[0]; S.p1 = G.par[1]; S.p2 = G.par[2];
    S.verts = t.kverts + G.kvert_off; S.nverts = G.kvert_cnt;
    S.cell_off = G.kmap_off >= 0 ? t.kcells + G.kmap_off : nullptr; S.cell_list = t.klists + G.klist_off;
    return S;
}

// h_S(d) for a direction d of any length `len` (plus the inflation radius), shape frame / placed by T
__device__ __forceinline__ float support_value(const ShapeRef &S, float3 d, float radius, float len) {
    const float3 sl = support_local(S, d);
    return dot3(d, sl) + radius * len;
}
__device__ __forceinline__ float support_value_placed(const ShapeRef &S, const float *T, float3 d, float radius, float len) {
    const float3 dl = make_float3(fmaf(T[0], d.x, fmaf(T[3], d.y, T[6] * d.z)), fmaf(T[1], d.x, fmaf(T[4], d.y, T[7] * d.z)),
                                  fmaf(T[2], d.x, fmaf(T[5], d.y, T[8] * d.z)));
    const float3 sl = support_local(S, dl);
    return dot3(d, se3_act(T, sl.x, sl.y, sl.z)) + radius * len;
}

// One out-of-line copy of the converged GJK serves the hull run, the core run and every triangle pair.
struct ConvOut {
    float dist;   // distance of the cores of the two shapes (0: they overlap); a lower bound when far != 0
    float3 v;     // closest point of A - B to the origin (A's frame): the separating direction is -v
    int far;      // certainly farther than `bound`
};
__device__ __noinline__ void gjk_converge(const ShapeRef &A, const ShapeRef &B, const float *T, float3 v0, float bound, ConvOut &o) {
    GjkState s;
    GjkOut r;
    gjk_init(s, v0);
    while (!gjk_step<true>(A, B, T, s, 0.f, bound, r)) {
    }
    o.dist = r.dist;
    o.far = r.far;
    o.v = s.v;
}

// ------------------------------------------------------------------------------------------
// Inner balls: a hit certificate that costs a handful of sphere tests
// ------------------------------------------------------------------------------------------
// DevGeom::balls lie inside the solid (inside the inner core of a surface geometry).  Two shapes with a pair of
// balls within the padding -- or a ball within the padding of a box, tested exactly in the box frame -- have
// cores within the padding: a hit, unless one shape may be enclosed by the other's surface (the caller asks
// the centre-line test then).  T: placement of b in a's frame.  slack: FP32 rounding of the centres (metres).
__device__ __forceinline__ bool inner_balls_hit(const DevGeom &GA, const DevGeom &GB, const float *T, float padding) {
    const float slack = 1e-5f;
    // all eight balls are read up front and every test is evaluated (unused balls have r = 0 and a test needs
    // r > 0): independent loads and straight-line arithmetic -- a loop with an early exit made a single-state
    // call wait for one cold table read after the other
    float4 ba[N_INNER_BALLS], bb[N_INNER_BALLS];
#pragma unroll
    for (int i = 0; i < N_INNER_BALLS; i++) {
        ba[i] = *reinterpret_cast<const float4 *>(GA.balls[i]);
        bb[i] = *reinterpret_cast<const float4 *>(GB.balls[i]);
    }
    bool hit = false;
    if (GA.shape == PBP_SHAPE_BOX) {
#pragma unroll
        for (int j = 0; j < N_INNER_BALLS; j++) {
            const float3 c = se3_act(T, bb[j].x, bb[j].y, bb[j].z);
            const float dx = fmaxf(fabsf(c.x) - GA.par[0], 0.f), dy = fmaxf(fabsf(c.y) - GA.par[1], 0.f), dz = fmaxf(fabsf(c.z) - GA.par[2], 0.f);
            const float r = bb[j].w + padding - slack;
            hit = hit || (bb[j].w > 0.f && r > 0.f && fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= r * r);
        }
        return hit;
    }
    if (GB.shape == PBP_SHAPE_BOX) {
#pragma unroll
        for (int i = 0; i < N_INNER_BALLS; i++) {
            const float3 c = se3_act_inv(T, make_float3(ba[i].x, ba[i].y, ba[i].z));
            const float dx = fmaxf(fabsf(c.x) - GB.par[0], 0.f), dy = fmaxf(fabsf(c.y) - GB.par[1], 0.f), dz = fmaxf(fabsf(c.z) - GB.par[2], 0.f);
            const float r = ba[i].w + padding - slack;
            hit = hit || (ba[i].w > 0.f && r > 0.f && fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= r * r);
        }
        return hit;
    }
#pragma unroll
    for (int j = 0; j < N_INNER_BALLS; j++) {
        const float3 cb = se3_act(T, bb[j].x, bb[j].y, bb[j].z);
#pragma unroll
        for (int i = 0; i < N_INNER_BALLS; i++) {
            const float dx = cb.x - ba[i].x, dy = cb.y - ba[i].y, dz = cb.z - ba[i].z;
            const float r = ba[i].w + bb[j].w + padding - slack;
            hit = hit || (ba[i].w > 0.f && bb[j].w > 0.f && r > 0.f && fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= r * r);
        }
    }
    return hit;
}

// ------------------------------------------------------------------------------------------
// Enclosure: is one shape of a core hit wholly inside the other's surface?
// ------------------------------------------------------------------------------------------
// `inner` pokes out of (or reaches) the enclosing hull  =>  the surfaces meet: with the cores overlapping
// and a vertex of the inner mesh outside the outer hull, the inner surface has points on both sides of
// the outer surface.  `inner` stays clear of every plane of the outer CORE by more than the padding  =>
// it lies inside the outer mesh with that clearance: free.  In between the triangles decide.
// pad_hit: with padding > 0 a point of the inner shape within (padding - tol) of a hull plane is within
// padding of the outer SURFACE (tol: two-sided distance between hull boundary and mesh); 0 otherwise.
__device__ __forceinline__ float enclosure_pad_hit(float padding, float tol_outer) { return padding > tol_outer ? padding - tol_outer : 0.f; }

enum { ENC_POKES = 0, ENC_ENCLOSED = 1, ENC_UNSURE = 2 };

// Centre-line test, one role: a direction in which the inner shape reaches at least as far as the outer one.
__device__ __forceinline__ bool pokes_along(const ShapeRef &in, const float *Tin, float rin, const ShapeRef &outer, float rout, float3 d, float pad_hit,
                                            bool in_is_placed) {
    const float len = sqrtf(dot3(d, d));
    const float hi = in_is_placed ? support_value_placed(in, Tin, d, rin, len) : support_value(in, d, rin, len);
    const float ho = in_is_placed ? support_value(outer, d, rout, len) : support_value_placed(outer, Tin, d, rout, len);
    return hi + pad_hit * len >= ho;
}

// Quick lane-level test of both flagged roles.  Returns true when every flagged role pokes out (a hit).
__device__ __noinline__ bool enclosure_quick_pokes(int fl, const DevGeom &GA, const DevGeom &GB, const ShapeRef &A, const ShapeRef &B, const float *T,
                                                   float padding) {
    const float3 ca = make_float3(GA.centre[0], GA.centre[1], GA.centre[2]);
    const float3 cb = se3_act(T, GB.centre[0], GB.centre[1], GB.centre[2]);
    float3 d = sub3(cb, ca);
    if (!(dot3(d, d) > 1e-12f)) d = make_float3(1.f, 0.f, 0.f);
    // b inside a's surface?  (b is placed by T, a is the frame)
    if ((fl & 1) && !pokes_along(B, T, GB.core_radius, A, GA.core_radius, d, enclosure_pad_hit(padding, GA.surface_tol), true)) return false;
    // a inside b's surface?
    if ((fl & 2) && !pokes_along(A, T, GA.core_radius, B, GB.core_radius, make_float3(-d.x, -d.y, -d.z), enclosure_pad_hit(padding, GB.surface_tol), false))
        return false;
    return true;
}

// Walk over the planes of the enclosing core, split over `nlanes` lanes.  T: placement of `inner` in the
// frame of G_outer.  max_faces limits the walk to the first (largest) HULL faces: then only ENC_POKES is
// a verdict.  gap_hull (optional): min over hull planes of the clearance (valid when not ENC_POKES).
__device__ __forceinline__ int enclosure_walk(const float4 *planes, const DevGeom &G_outer, const ShapeRef &inner, float inner_radius, const float *T,
                                              float padding, int lane, int nlanes, int max_faces, float *gap_hull) {
    const int nhull = G_outer.plane_hull_cnt;
    const int nfaces = max_faces < nhull ? max_faces : G_outer.plane_cnt;
    const float pad_hit = enclosure_pad_hit(padding, G_outer.surface_tol);
    float worst_h = -FLT_MAX, worst_k = -FLT_MAX;
    // two faces per lane and trip: every face is a chain of dependent table reads (plane -> support-map cell ->
    // candidate list -> vertices); the second face's chain runs in the shadow of the first
    for (int f0 = 0; f0 < nfaces; f0 += 2 * nlanes) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int f = f0 + u * nlanes + lane;
            if (f < nfaces) {
                const float4 pl = planes[G_outer.plane_off + f];
                const float3 dl = make_float3(fmaf(T[0], pl.x, fmaf(T[3], pl.y, T[6] * pl.z)), fmaf(T[1], pl.x, fmaf(T[4], pl.y, T[7] * pl.z)),
                                              fmaf(T[2], pl.x, fmaf(T[5], pl.y, T[8] * pl.z)));
                const float3 sl = support_local(inner, dl);
                const float3 sw = se3_act(T, sl.x, sl.y, sl.z);
                const float ex = fmaf(pl.x, sw.x, fmaf(pl.y, sw.y, pl.z * sw.z)) + inner_radius - pl.w;
                worst_k = fmaxf(worst_k, ex);
                if (f < nhull) worst_h = fmaxf(worst_h, ex);
            }
        }
        // one hull face the inner shape reaches settles it; nearly all parked items end in the first batch
        const bool out = worst_h >= -pad_hit;
        if (nlanes > 1 ? __any_sync(0xffffffffu, out) : out) return ENC_POKES;
    }
    if (nlanes > 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            worst_h = fmaxf(worst_h, __shfl_xor_sync(0xffffffffu, worst_h, o));
            worst_k = fmaxf(worst_k, __shfl_xor_sync(0xffffffffu, worst_k, o));
        }
    }
    if (gap_hull) *gap_hull = -worst_h;
    if (max_faces < nhull) return ENC_UNSURE;
    return -worst_k > padding ? ENC_ENCLOSED : ENC_UNSURE;
}

// Both flagged roles (bit 0: b inside a's surface, bit 1: a inside b's).  T: placement of b in a's frame.
// ENC_POKES: neither shape is enclosed (the core hit stands); ENC_ENCLOSED: one is, with clearance beyond
// the padding; ENC_UNSURE: the triangles decide.  gap_hull: clearance to the hull planes of the enclosing role.
__device__ __forceinline__ int enclosure_roles(const float4 *planes, int fl, const DevGeom &GA, const DevGeom &GB, const ShapeRef &A, const ShapeRef &B,
                                               const float *T, float padding, int lane, int nlanes, int max_faces, float *gap_hull, float *tol_outer) {
    int res = ENC_POKES;
    if (fl & 1) {
        float g = 0.f;
        const int r = enclosure_walk(planes, GA, B, GB.core_radius, T, padding, lane, nlanes, max_faces, &g);
        if (r == ENC_ENCLOSED) { if (gap_hull) *gap_hull = g; if (tol_outer) *tol_outer = GA.surface_tol; return ENC_ENCLOSED; }
        if (r == ENC_UNSURE) { res = ENC_UNSURE; if (gap_hull) *gap_hull = g; if (tol_outer) *tol_outer = GA.surface_tol; }
    }
    if (fl & 2) {
        float Ti[12];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Ti[3 * i + 0] = T[i]; Ti[3 * i + 1] = T[3 + i]; Ti[3 * i + 2] = T[6 + i];
            Ti[9 + i] = -(T[i] * T[9] + T[3 + i] * T[10] + T[6 + i] * T[11]);
        }
        float g = 0.f;
        const int r = enclosure_walk(planes, GB, A, GA.core_radius, Ti, padding, lane, nlanes, max_faces, &g);
        if (r == ENC_ENCLOSED) { if (gap_hull) *gap_hull = g; if (tol_outer) *tol_outer = GB.surface_tol; return ENC_ENCLOSED; }
        if (r == ENC_UNSURE) { res = ENC_UNSURE; if (gap_hull) *gap_hull = g; if (tol_outer) *tol_outer = GB.surface_tol; }
    }
    return res;
}

// ------------------------------------------------------------------------------------------
// Distance between two triangles, closed form
// ------------------------------------------------------------------------------------------
// The minimum distance of two triangles is attained between an edge of one and an edge of the other, or between
// a vertex of one and the face of the other -- unless they intersect, which is the case exactly when an edge of
// one passes through the face of the other.  Straight-line FP32 arithmetic (about 900 flops), so the 32 lanes of
// a warp stay converged over their 32 triangle pairs; a converged GJK run on the same pair is ~20 dependent,
// divergent iterations.

// squared distance between segments [p1, q1] and [p2, q2] (Ericson, Real-Time Collision Detection 5.1.9)
__device__ __noinline__ float segseg_dist2(float3 p1, float3 q1, float3 p2, float3 q2) {
    const float3 d1 = sub3(q1, p1), d2 = sub3(q2, p2), r = sub3(p1, p2);
    const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    float sp = 0.f, tp = 0.f;
    if (a > 0.f && e > 0.f) {
        const float b = dot3(d1, d2), c = dot3(d1, r);
        const float den = a * e - b * b;
        sp = den > 1e-12f * a * e ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
        tp = (b * sp + f) / e;
        if (tp < 0.f) { tp = 0.f; sp = fminf(fmaxf(-c / a, 0.f), 1.f); }
        else if (tp > 1.f) { tp = 1.f; sp = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
    } else if (a > 0.f) {
        sp = fminf(fmaxf(-dot3(d1, r) / a, 0.f), 1.f);
    } else if (e > 0.f) {
        tp = fminf(fmaxf(f / e, 0.f), 1.f);
    }
    const float3 w = sub3(madd3(p1, d1, sp), madd3(p2, d2, tp));
    return dot3(w, w);
}

// squared distance between the point p and the triangle (a, b, c), counted only when the closest point lies in the
// triangle's INTERIOR or on it along the normal (edges and vertices are covered by the segment tests): FLT_MAX else
__device__ __noinline__ float point_face_dist2(float3 p, float3 a, float3 b, float3 c) {
    const float3 ab = sub3(b, a), ac = sub3(c, a), ap = sub3(p, a);
    const float3 n = cross3(ab, ac);
    const float nn = dot3(n, n);
    if (!(nn > 0.f)) return FLT_MAX;
    // barycentric coordinates of the projection
    const float v = dot3(cross3(ap, ac), n), w = dot3(cross3(ab, ap), n);
    if (v < 0.f || w < 0.f || v + w > nn) return FLT_MAX;
    const float h = dot3(ap, n);
    return h * h / nn;
}

// does the segment [p, q] pass through the triangle (a, b, c)?  (p and q strictly on different sides, or touching)
__device__ __noinline__ bool segment_through_face(float3 p, float3 q, float3 a, float3 b, float3 c) {
    const float3 ab = sub3(b, a), ac = sub3(c, a);
    const float3 n = cross3(ab, ac);
    const float dp = dot3(sub3(p, a), n), dq = dot3(sub3(q, a), n);
    if ((dp > 0.f && dq > 0.f) || (dp < 0.f && dq < 0.f) || dp == dq) return false;
    const float t = dp / (dp - dq);
    const float3 x = madd3(p, sub3(q, p), t);
    const float3 ax = sub3(x, a);
    const float nn = dot3(n, n);
    const float v = dot3(cross3(ax, ac), n), w = dot3(cross3(ab, ax), n);
    return v >= 0.f && w >= 0.f && v + w <= nn;
}

// (Loops kept rolled around out-of-line helpers on purpose: a warp runs this once per item, and 1500 straight-line
//  instructions fetched cold from the instruction cache took 25 k cycles; the rolled form re-executes ~300.)
__device__ __noinline__ float tri_tri_distance(const float3 *A, const float3 *B) {
    float d2 = FLT_MAX;
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        const float3 a0 = A[i], a1 = A[i == 2 ? 0 : i + 1];
#pragma unroll 1
        for (int j = 0; j < 3; j++) d2 = fminf(d2, segseg_dist2(a0, a1, B[j], B[j == 2 ? 0 : j + 1]));
    }
    bool cut = false;
#pragma unroll 1
    for (int i = 0; i < 3; i++) {
        d2 = fminf(d2, point_face_dist2(A[i], B[0], B[1], B[2]));
        d2 = fminf(d2, point_face_dist2(B[i], A[0], A[1], A[2]));
        cut = cut || segment_through_face(A[i], A[i == 2 ? 0 : i + 1], B[0], B[1], B[2]);
        cut = cut || segment_through_face(B[i], B[i == 2 ? 0 : i + 1], A[0], A[1], A[2]);
    }
    return cut ? 0.f : sqrtf(d2);
}

// ------------------------------------------------------------------------------------------
// Triangle stage (one warp per item)
// ------------------------------------------------------------------------------------------
constexpr int TRI_LIST_CAP = 384;   // triangles of one geometry the warp's lists hold per pass

// The triangles [t_begin, t_end) of one geometry that can come within `reach` of the ball around c_other and,
// with a separating direction n, lie on the near side of the slab (SIDE_TOP: max of n.x >= lim; else min of
// n.x <= lim): their indices go to `list`, in order; returns how many.  PLACED: the vertices are moved by T first.
// Four triangles per lane are in flight at a time: the pass is a chain of table reads (L2 latency), not arithmetic.
template <bool PLACED, bool SIDE_TOP>
__device__ __forceinline__ int cull_triangles(const float4 *tverts, int t_begin, int t_end, const float *T, float3 c_other, float reach, bool have_n,
                                              float3 n, float lim, uint16_t *list, int lane) {
    constexpr int UNR = PBP_LANES > 1 ? 4 : 1;
    const unsigned lt_mask = (1u << lane) - 1u;
    int count = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += UNR * PBP_LANES) {
        float4 v[UNR][3];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int t = min(t0 + u * PBP_LANES + lane, t_end - 1);
            const float4 *tv = tverts + 3 * (size_t)t;
            v[u][0] = tv[0]; v[u][1] = tv[1]; v[u][2] = tv[2];
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int t = t0 + u * PBP_LANES + lane;
            float3 p0 = make_float3(v[u][0].x, v[u][0].y, v[u][0].z), p1 = make_float3(v[u][1].x, v[u][1].y, v[u][1].z),
                   p2 = make_float3(v[u][2].x, v[u][2].y, v[u][2].z);
            if (PLACED) { p0 = se3_act(T, p0.x, p0.y, p0.z); p1 = se3_act(T, p1.x, p1.y, p1.z); p2 = se3_act(T, p2.x, p2.y, p2.z); }
            const float gx = fmaxf(fmaxf(fminf(fminf(p0.x, p1.x), p2.x) - c_other.x, c_other.x - fmaxf(fmaxf(p0.x, p1.x), p2.x)), 0.f);
            const float gy = fmaxf(fmaxf(fminf(fminf(p0.y, p1.y), p2.y) - c_other.y, c_other.y - fmaxf(fmaxf(p0.y, p1.y), p2.y)), 0.f);
            const float gz = fmaxf(fmaxf(fminf(fminf(p0.z, p1.z), p2.z) - c_other.z, c_other.z - fmaxf(fmaxf(p0.z, p1.z), p2.z)), 0.f);
            bool keep = t < t_end && fmaf(gx, gx, fmaf(gy, gy, gz * gz)) <= reach * reach;
            if (have_n) {
                const float d0 = dot3(n, p0), d1 = dot3(n, p1), d2 = dot3(n, p2);
                keep = keep && (SIDE_TOP ? fmaxf(fmaxf(d0, d1), d2) >= lim : fminf(fminf(d0, d1), d2) <= lim);
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (keep) list[count + __popc(km & lt_mask)] = (uint16_t)t;
            count += __popc(km);
        }
    }
    return count;
}

// Smallest distance between the (cores of the) two geometries over all triangle pairs within U, FLT_MAX when
// there is none.  A geometry without triangles is its solid (primitive or hull) as one piece.  n (unit, A's
// frame, pointing from A to B; have_n): only triangles within the slab the other shape can reach along n
// can be within U; triangles farther than U from the other shape's bounding ball never are.
// first_hit: return as soon as some pair is within U (boolean queries).  la / lb: per-warp shared scratch.
__device__ __noinline__ float triangle_stage(const SurfTables &tab, const DevGeom &GA, const DevGeom &GB, const float *T, float3 n, bool have_n, float U,
                                             bool first_hit, uint16_t *la, uint16_t *lb, int lane) {
    const int nta = GA.tri_cnt, ntb = GB.tri_cnt;
    const ShapeRef AH = hull_shape(tab, GA), BH = hull_shape(tab, GB);
    const float3 ca = make_float3(GA.centre[0], GA.centre[1], GA.centre[2]);
    const float3 cb = se3_act(T, GB.centre[0], GB.centre[1], GB.centre[2]);
    const float ra = GA.r_out - GA.core_radius, rb = GB.r_out - GB.core_radius;   // balls around the CORES
    const float slack = 2e-6f;
    float hi_a = FLT_MAX, lo_b = -FLT_MAX;
    if (have_n) {
        hi_a = support_value(AH, n, 0.f, 1.f);                                                // max over A of n.x
        lo_b = -support_value_placed(BH, T, make_float3(-n.x, -n.y, -n.z), 0.f, 1.f);         // min over B of n.x
    }
    float best = FLT_MAX;
    float bound = U;
    const long long tq0 = DBG_CLOCK();
    for (int a0 = 0; a0 < (nta > 0 ? nta : 1); a0 += TRI_LIST_CAP) {
        // ---- A's triangles that can matter
        int ca_n = 0;
        if (nta == 0) {
            ca_n = 1;
        } else {
            const int a1 = min(a0 + TRI_LIST_CAP, nta);
            ca_n = cull_triangles<false, true>(tab.tverts + 3 * (size_t)GA.tri_off, a0, a1, T, cb, rb + U + slack, have_n, n, lo_b - U - slack, la, lane);
        }
        if (lane == 0) DBG_ADD(40, DBG_CLOCK() - tq0);
        if (ca_n == 0) continue;
        for (int b0 = 0; b0 < (ntb > 0 ? ntb : 1); b0 += TRI_LIST_CAP) {
            // ---- B's triangles that can matter (tested in A's frame)
            int cb_n = 0;
            if (ntb == 0) {
                cb_n = 1;
            } else {
                const int b1 = min(b0 + TRI_LIST_CAP, ntb);
                cb_n = cull_triangles<true, false>(tab.tverts + 3 * (size_t)GB.tri_off, b0, b1, T, ca, ra + U + slack, have_n, n, hi_a + U + slack, lb, lane);
            }
            if (lane == 0) DBG_ADD(41, DBG_CLOCK() - tq0);
            if (cb_n == 0) continue;
            __syncwarp();
            // ---- every pair of the two lists, the lanes sharing them
            const int npairs = ca_n * cb_n;
            if (lane == 0) { DBG_ADD(7, npairs); DBG_MAX(8, npairs); DBG_ADD(15, ca_n); DBG_ADD(16, cb_n); }
            for (int i0 = 0; i0 < npairs; i0 += PBP_LANES) {
                const int i = i0 + lane;
                if (i < npairs) {
                    const int ia = i / cb_n, ib = i - ia * cb_n;
                    float4 va[4], vb[4];
                    ShapeRef SA = AH, SB = BH;
                    float3 v0a = ca, v0b = cb;
                    if (nta > 0) {
                        const float4 *tv = tab.tverts + 3 * (size_t)(GA.tri_off + la[ia]);
                        va[0] = tv[0]; va[1] = tv[1]; va[2] = tv[2];
                        va[3] = va[2];
                        SA.shape = PBP_SHAPE_CONVEX; SA.verts = va; SA.nverts = 4; SA.cell_off = nullptr; SA.cell_list = nullptr;
                        v0a = make_float3((va[0].x + va[1].x + va[2].x) * (1.f / 3.f), (va[0].y + va[1].y + va[2].y) * (1.f / 3.f),
                                          (va[0].z + va[1].z + va[2].z) * (1.f / 3.f));
                    }
                    if (ntb > 0) {
                        const float4 *tv = tab.tverts + 3 * (size_t)(GB.tri_off + lb[ib]);
                        vb[0] = tv[0]; vb[1] = tv[1]; vb[2] = tv[2];
                        vb[3] = vb[2];
                        SB.shape = PBP_SHAPE_CONVEX; SB.verts = vb; SB.nverts = 4; SB.cell_off = nullptr; SB.cell_list = nullptr;
                        v0b = se3_act(T, (vb[0].x + vb[1].x + vb[2].x) * (1.f / 3.f), (vb[0].y + vb[1].y + vb[2].y) * (1.f / 3.f),
                                      (vb[0].z + vb[1].z + vb[2].z) * (1.f / 3.f));
                    }
                    if (nta > 0 && ntb > 0) {
                        const float3 ta[3] = {make_float3(va[0].x, va[0].y, va[0].z), make_float3(va[1].x, va[1].y, va[1].z), make_float3(va[2].x, va[2].y, va[2].z)};
                        const float3 tb[3] = {se3_act(T, vb[0].x, vb[0].y, vb[0].z), se3_act(T, vb[1].x, vb[1].y, vb[1].z), se3_act(T, vb[2].x, vb[2].y, vb[2].z)};
                        // boxes of the two triangles first: a non-convex mesh (UR upper arm: 1176 triangles) leaves
                        // thousands of pairs per item, nearly all of them far apart
                        const float gx = fmaxf(fmaxf(fminf(fminf(ta[0].x, ta[1].x), ta[2].x) - fmaxf(fmaxf(tb[0].x, tb[1].x), tb[2].x),
                                                     fminf(fminf(tb[0].x, tb[1].x), tb[2].x) - fmaxf(fmaxf(ta[0].x, ta[1].x), ta[2].x)), 0.f);
                        const float gy = fmaxf(fmaxf(fminf(fminf(ta[0].y, ta[1].y), ta[2].y) - fmaxf(fmaxf(tb[0].y, tb[1].y), tb[2].y),
                                                     fminf(fminf(tb[0].y, tb[1].y), tb[2].y) - fmaxf(fmaxf(ta[0].y, ta[1].y), ta[2].y)), 0.f);
                        const float gz = fmaxf(fmaxf(fminf(fminf(ta[0].z, ta[1].z), ta[2].z) - fmaxf(fmaxf(tb[0].z, tb[1].z), tb[2].z),
                                                     fminf(fminf(tb[0].z, tb[1].z), tb[2].z) - fmaxf(fmaxf(ta[0].z, ta[1].z), ta[2].z)), 0.f);
                        const float lim = fminf(bound, best) + 2e-6f;
                        if (fmaf(gx, gx, fmaf(gy, gy, gz * gz)) <= lim * lim) {
                            const float d = tri_tri_distance(ta, tb);
                            if (d <= bound && d < best) best = d;
                        }
                    } else {
                        ConvOut o;
                        gjk_converge(SA, SB, T, sub3(v0a, v0b), bound, o);
                        if (!o.far && o.dist < best) best = o.dist;
                    }
                }
                // tighten the bound with what the warp has found so far
                float wb = best;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wb = fminf(wb, __shfl_xor_sync(0xffffffffu, wb, o));
                if (wb <= U && first_hit) { if (lane == 0) DBG_ADD(42, DBG_CLOCK() - tq0); return wb; }
                bound = fminf(bound, wb);
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    return best <= U ? best : FLT_MAX;
}

// ------------------------------------------------------------------------------------------
// Lane level: one parked item
// ------------------------------------------------------------------------------------------
enum { ST_NONE = 0, ST_HIT = 1, ST_WALK = 2, ST_TRIANGLES = 3, ST_VALUE = 4 };
#ifndef PBP_SR_HULL_ITERS
#define PBP_SR_HULL_ITERS 8   // iterations an undecided item's hull run gets before the triangles decide
#endif
struct LaneVerdict {
    int status;     // ST_NONE: free / cannot improve the minimum; ST_HIT: the pair collides; ST_WALK: the warp walks the
                    // enclosing core's planes; ST_TRIANGLES: the warp runs the triangle stage; ST_VALUE: dist is exact
    float dist;     // ST_VALUE: the pair's core distance
    float U;        // ST_TRIANGLES: pairs farther apart than this cannot matter (core distance)
    float3 n;       // ST_TRIANGLES: separating direction of the cores (unit, A's frame), valid when have_n
    int have_n;
};

// Boolean queries.  parked: core hit of a flagged pair; otherwise the item converged between margin and cut.
// v0: where a previous GJK run on the cores of this item converged (zero vector: unknown) -- a warm start: the
// hull run then ends after one support call when the hulls are apart along it, and both runs converge in a few.
__device__ __forceinline__ LaneVerdict decide_boolean(const SurfTables &tab, int fl, bool parked, const DevGeom &GA, const DevGeom &GB, const float *T,
                                                      float padding, float3 v0 = make_float3(0.f, 0.f, 0.f)) {
    LaneVerdict r;
    r.status = ST_NONE; r.dist = 0.f; r.U = 0.f; r.n = make_float3(0.f, 0.f, 0.f); r.have_n = 0;
    const ShapeRef AH = hull_shape(tab, GA), BH = hull_shape(tab, GB);
    const float margin = padding + GA.core_radius + GB.core_radius;
    if (parked) {
        // the centre line settles most parked items, the enclosing hull's 8 largest faces most of the rest
        if (enclosure_quick_pokes(fl, GA, GB, AH, BH, T, padding) ||
            enclosure_roles(tab.planes, fl, GA, GB, AH, BH, T, padding, 0, 1, 8, nullptr, nullptr) == ENC_POKES)
            r.status = ST_HIT;
        else
            r.status = ST_WALK;
        return r;
    }
    ConvOut h;
    const bool warm = dot3(v0, v0) > 1e-30f;   // (any non-zero vector: cores 1e-7 m apart are as undecided as cores 1e-4 m apart)
    const float3 start = warm ? v0 : gjk_start_direction(GA, GB, T);
    if (warm && !(fl & PAIR_NOCORE)) {
        // The cores are known to be farther apart than the margin (|v0| > margin): all that is asked of the hulls
        // is "within the margin or not", which the boolean run answers without converging -- apart: free; within:
        // the triangles decide, culled along the cores' separating direction.
        GjkState s;
        GjkOut g;
        gjk_init(s, start);
        bool settled = true;
        while (!gjk_step<false>(AH, BH, T, s, margin, margin, g)) {
            // hulls about `margin` apart converge slowly in FP32 (up to 48 iterations): such a lane was the tail of
            // the whole launch -- after a few iterations the (exact) triangle stage takes the item as it is
            if (s.it >= PBP_SR_HULL_ITERS) { settled = false; break; }
        }
        if (settled && (g.far || !g.hit)) return r;
        r.status = ST_TRIANGLES;
        r.U = margin;
        const float inv = rsqrtf(dot3(v0, v0));
        r.n = make_float3(-v0.x * inv, -v0.y * inv, -v0.z * inv);
        r.have_n = 1;
        return r;
    }
    gjk_converge(AH, BH, T, start, margin, h);
    if (h.far || h.dist > margin) return r;                 // the hulls are apart: free
    ConvOut k = h;
    if (!(fl & PAIR_NOCORE)) {
        if (warm) {
            // v0 IS the converged closest point of the cores' difference: nothing to run again
            k.v = v0; k.dist = sqrtf(dot3(v0, v0)); k.far = 0;
        } else {
            const ShapeRef AK = core_shape(tab, GA), BK = core_shape(tab, GB);
            gjk_converge(AK, BK, T, start, FLT_MAX, k);
        }
        if (k.dist - h.dist <= 1e-6f) { r.status = ST_HIT; return r; }   // both bounds agree: within the margin
    }
    r.status = ST_TRIANGLES;
    r.U = margin;
    const float vv = dot3(k.v, k.v);
    if (k.dist > 0.f && vv > 1e-20f) {
        const float inv = rsqrtf(vv);
        r.n = make_float3(-k.v.x * inv, -k.v.y * inv, -k.v.z * inv);
        r.have_n = 1;
    }
    return r;
}

// Distance queries.  cur: the state's current upper bound of the minimum (surface distance).
__device__ __forceinline__ LaneVerdict decide_distance(const SurfTables &tab, int fl, const DevGeom &GA, const DevGeom &GB, const float *T, float cur) {
    LaneVerdict r;
    r.status = ST_NONE; r.dist = 0.f; r.U = 0.f; r.n = make_float3(0.f, 0.f, 0.f); r.have_n = 0;
    const ShapeRef AH = hull_shape(tab, GA), BH = hull_shape(tab, GB);
    const float radii = GA.core_radius + GB.core_radius;
    ConvOut h;
    gjk_converge(AH, BH, T, gjk_start_direction(GA, GB, T), cur + radii, h);
    if (h.far || fmaxf(h.dist - radii, 0.f) > cur) return r;   // the surfaces are at least as far apart as the hulls
    if (fl & PAIR_NOCORE) {
        // no upper bound from a core: the triangles within the state's current bound decide
        r.status = ST_TRIANGLES;
        r.U = cur + radii + 1e-6f;
        const float vv = dot3(h.v, h.v);
        if (h.dist > 0.f && vv > 1e-20f) {
            const float inv = rsqrtf(vv);
            r.n = make_float3(-h.v.x * inv, -h.v.y * inv, -h.v.z * inv);
            r.have_n = 1;
        }
        return r;
    }
    const ShapeRef AK = core_shape(tab, GA), BK = core_shape(tab, GB);
    ConvOut k;
    gjk_converge(AK, BK, T, gjk_start_direction(GA, GB, T), FLT_MAX, k);
    if (k.dist > 0.f) {
        if (k.dist - h.dist <= 2e-6f) { r.status = ST_VALUE; r.dist = h.dist; return r; }
        r.status = ST_TRIANGLES;
        r.U = fminf(k.dist, cur + radii) + 1e-6f;
        const float vv = dot3(k.v, k.v);
        if (vv > 1e-20f) {
            const float inv = rsqrtf(vv);
            r.n = make_float3(-k.v.x * inv, -k.v.y * inv, -k.v.z * inv);
            r.have_n = 1;
        }
        // d(surfaces) <= d(cores) needs the segment between the cores' closest points to cross both surfaces, which
        // it does not when one shape sits INSIDE the other's solid without touching its core (a small box in the
        // shell between a UR link's surface and its inner cell).  Overlapping hulls of a pair with an enclosure flag
        // whose would-be inner shape does not poke out: the warp measures its clearance to the enclosing hull's
        // planes and searches that far instead (ST_WALK with U > 0: "the cores are apart").
        if ((fl & PAIR_ENCLOSE) && h.dist <= 0.f && !enclosure_quick_pokes(fl, GA, GB, AH, BH, T, 0.f)) r.status = ST_WALK;
        return r;
    }
    // the cores overlap: the surfaces meet unless one shape is enclosed by the other's
    if (!(fl & PAIR_ENCLOSE) || enclosure_quick_pokes(fl, GA, GB, AH, BH, T, 0.f)) { r.status = ST_VALUE; r.dist = 0.f; return r; }
    r.status = ST_WALK;
    return r;
}

// ------------------------------------------------------------------------------------------
// Warp level: settle what the lanes left open
// ------------------------------------------------------------------------------------------
// Warp-level part shared by k_surface_refine and k_small_check.  Every lane arrives with its own item
// (pair p, output group, placement T, lane verdict v); items whose verdict needs the warp (the walk over
// the enclosing core's planes, the triangle stage) are taken one after the other, their data broadcast
// from the owning lane.  on_hit(p, group, aux) / on_value(p, group, aux, surface distance) publish; they
// are called by lane 0 only.  decided(group): skip an item whose group has been decided meanwhile
// (boolean queries that stop at the first hit); evaluated by lane 0.
template <int MODE, typename FHit, typename FValue, typename FDecided>
__device__ __forceinline__ void settle_warp_items(const DevModel &M, const SurfTables &tab, const LaneVerdict &v, int p, int32_t group, int64_t aux,
                                                  const float *T, float padding, uint16_t *la, uint16_t *lb, int lane, FHit on_hit, FValue on_value,
                                                  FDecided decided) {
    unsigned todo = __ballot_sync(0xffffffffu, v.status == ST_WALK || v.status == ST_TRIANGLES);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        float Ts[12];
#pragma unroll
        for (int k = 0; k < 12; k++) Ts[k] = __shfl_sync(0xffffffffu, T[k], src);
        const int ps = __shfl_sync(0xffffffffu, p, src);
        const int32_t gs = __shfl_sync(0xffffffffu, group, src);
        const int64_t as = __shfl_sync(0xffffffffu, aux, src);
        int st = __shfl_sync(0xffffffffu, v.status, src);
        float U = __shfl_sync(0xffffffffu, v.U, src);
        float3 n = make_float3(__shfl_sync(0xffffffffu, v.n.x, src), __shfl_sync(0xffffffffu, v.n.y, src), __shfl_sync(0xffffffffu, v.n.z, src));
        int have_n = __shfl_sync(0xffffffffu, v.have_n, src);
        int skip = lane == 0 ? (decided(gs) ? 1 : 0) : 0;      // one lane reads, everyone follows
        skip = __shfl_sync(0xffffffffu, skip, 0);
        if (skip) continue;
        const int fl = M.pair_enclose[ps];
        const BpPair pr = M.bppairs[ps];
        const DevGeom &GA = M.geoms[pr.a];
        const DevGeom &GB = M.geoms[pr.b];
        const float radii = GA.core_radius + GB.core_radius;
        if (st == ST_WALK) {
            const ShapeRef AH = hull_shape(tab, GA), BH = hull_shape(tab, GB);
            float gap = 0.f, tolo = 0.f;
            const long long tw0 = DBG_CLOCK();
            const int res = enclosure_roles(tab.planes, fl, GA, GB, AH, BH, Ts, MODE == MODE_DIST ? 0.f : padding, lane, PBP_LANES, INT_MAX, &gap, &tolo);
            if (lane == 0) { DBG_ADD(12, DBG_CLOCK() - tw0); DBG_ADD(17 + res, 1); }
            const bool cores_apart = MODE == MODE_DIST && U > 0.f;      // (see decide_distance)
            if (res == ENC_POKES && !cores_apart) {
                if (lane == 0) { if (MODE == MODE_DIST) on_value(ps, gs, as, 0.f); else on_hit(ps, gs, as); }
                continue;
            }
            if (MODE != MODE_DIST && res == ENC_ENCLOSED) continue;     // inside the other's surface, clear of it: free
            st = ST_TRIANGLES;
            if (!(res == ENC_POKES && cores_apart)) {
                // the triangles decide: no separating direction here, the bounding balls do the culling
                have_n = 0;
                U = MODE == MODE_DIST ? fmaxf(gap, 0.f) + tolo + radii + 1e-6f : padding + radii;
            }   // (pokes out with the cores apart: the lane's bound and direction stand)
        }
        const long long tt0 = DBG_CLOCK();
        const float d = triangle_stage(tab, GA, GB, Ts, n, have_n != 0, U, MODE != MODE_DIST, la, lb, lane);
        if (lane == 0) { const long long dt = DBG_CLOCK() - tt0; DBG_ADD(6, 1); DBG_ADD(9, dt); DBG_MAX(10, dt); DBG_ADD(20, have_n ? 1 : 0); DBG_ADD(21, d <= U ? 1 : 0); }
        if (lane == 0 && d <= U) {
            if (MODE == MODE_DIST) on_value(ps, gs, as, fmaxf(d - radii, 0.f)); else on_hit(ps, gs, as);
        }
    }
}

}  // namespace pbp
