// FP32 GJK on support functions (sphere / box / cylinder / capsule cores, convex hulls).
//
// One thread runs one (state, pair) item; all lanes of a warp work on the SAME pair, so the
// hull vertex loop of the support function reads identical shared-memory addresses in every
// lane (hardware broadcast, one wavefront per LDS.128, no bank conflicts) and never diverges.
// Spheres and capsules are handled as core (point / segment) + radius folded into `margin`.
#pragma once
#include <float.h>

#include "pbp_model.cuh"

#if !defined(__CUDACC__) && !defined(__noinline__)
#define __noinline__ __attribute__((noinline))   // host build of this header (tests/host_emu)
#endif

namespace pbp {

struct ShapeRef {
    int shape;
    float p0, p1, p2;       // box: half sides | cylinder/capsule: p0 = radius, p1 = half length
    const float4 *verts;    // convex: padded vertex table (shared memory or global)
    int nverts;             // multiple of 4
    const uint16_t *cell_off;  // convex: support map (pbp_supmap.h), NULL -> scan every vertex
    const uint16_t *cell_list;
};

#ifndef PBP_SUPMAP_K
#define PBP_SUPMAP_K 16
#endif
constexpr int SUPMAP_KD = PBP_SUPMAP_K;  // must equal SUPMAP_K of pbp_supmap.h

__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ float3 madd3(float3 a, float3 d, float t) {
    return make_float3(fmaf(d.x, t, a.x), fmaf(d.y, t, a.y), fmaf(d.z, t, a.z));
}

// argmax_i <v_i, d> over a padded vertex table; 4 independent running maxima for ILP.
__device__ __noinline__ float3 support_convex(const float4 *__restrict__ v, int n, float3 d) {
    float b0 = -FLT_MAX, b1 = -FLT_MAX, b2 = -FLT_MAX, b3 = -FLT_MAX;
    int i0 = 0, i1 = 1, i2 = 2, i3 = 3;
#pragma unroll 2
    for (int i = 0; i < n; i += 4) {
        const float4 p0 = v[i], p1 = v[i + 1], p2 = v[i + 2], p3 = v[i + 3];
        const float d0 = fmaf(p0.x, d.x, fmaf(p0.y, d.y, p0.z * d.z));
        const float d1 = fmaf(p1.x, d.x, fmaf(p1.y, d.y, p1.z * d.z));
        const float d2 = fmaf(p2.x, d.x, fmaf(p2.y, d.y, p2.z * d.z));
        const float d3 = fmaf(p3.x, d.x, fmaf(p3.y, d.y, p3.z * d.z));
        if (d0 > b0) { b0 = d0; i0 = i; }
        if (d1 > b1) { b1 = d1; i1 = i + 1; }
        if (d2 > b2) { b2 = d2; i2 = i + 2; }
        if (d3 > b3) { b3 = d3; i3 = i + 3; }
    }
    if (b1 > b0) { b0 = b1; i0 = i1; }
    if (b3 > b2) { b2 = b3; i2 = i3; }
    if (b2 > b0) { i0 = i2; }
    const float4 p = v[i0];
    return make_float3(p.x, p.y, p.z);
}

// Exact support through the hull's support map: the cell of the direction (cube-face chart) lists
// every vertex that can be the maximiser for a direction in that cell (~3 instead of ~150).
__device__ __forceinline__ float3 support_convex_map(const ShapeRef &s, float3 d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int a;
    float da, db, dc;
    if (ax >= ay && ax >= az) { a = 0; da = d.x; db = d.y; dc = d.z; }
    else if (ay >= az) { a = 1; da = d.y; db = d.z; dc = d.x; }
    else { a = 2; da = d.z; db = d.x; dc = d.y; }
    const float inv = 1.0f / fabsf(da);
    const int iu = min(SUPMAP_KD - 1, max(0, (int)((db * inv + 1.0f) * (0.5f * SUPMAP_KD))));
    const int iw = min(SUPMAP_KD - 1, max(0, (int)((dc * inv + 1.0f) * (0.5f * SUPMAP_KD))));
    const int cell = (2 * a + (da < 0.f ? 1 : 0)) * (SUPMAP_KD * SUPMAP_KD) + iu * SUPMAP_KD + iw;
    const int off = s.cell_off[cell], end = s.cell_off[cell + 1];
    float best = -FLT_MAX;
    float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
    // four candidates per trip: the index loads, then the vertex loads, are independent of each
    // other, so their shared-memory latencies overlap (lists are padded by 4 entries at the end)
    for (int j = off; j < end; j += 4) {
        const int i0 = s.cell_list[j], i1 = s.cell_list[j + 1], i2 = s.cell_list[j + 2], i3 = s.cell_list[j + 3];
        const float4 p0 = s.verts[i0], p1 = s.verts[i1], p2 = s.verts[i2], p3 = s.verts[i3];
        const float d0 = fmaf(p0.x, d.x, fmaf(p0.y, d.y, p0.z * d.z));
        const float d1 = j + 1 < end ? fmaf(p1.x, d.x, fmaf(p1.y, d.y, p1.z * d.z)) : -FLT_MAX;
        const float d2 = j + 2 < end ? fmaf(p2.x, d.x, fmaf(p2.y, d.y, p2.z * d.z)) : -FLT_MAX;
        const float d3 = j + 3 < end ? fmaf(p3.x, d.x, fmaf(p3.y, d.y, p3.z * d.z)) : -FLT_MAX;
        if (d0 > best) { best = d0; bp = p0; }
        if (d1 > best) { best = d1; bp = p1; }
        if (d2 > best) { best = d2; bp = p2; }
        if (d3 > best) { best = d3; bp = p3; }
    }
    return make_float3(bp.x, bp.y, bp.z);
}

__device__ __forceinline__ float3 support_local(const ShapeRef &s, float3 d) {
    switch (s.shape) {
    case PBP_SHAPE_SPHERE:
        return make_float3(0.f, 0.f, 0.f);
    case PBP_SHAPE_CAPSULE:
        return make_float3(0.f, 0.f, d.z >= 0.f ? s.p1 : -s.p1);
    case PBP_SHAPE_BOX:
        return make_float3(d.x >= 0.f ? s.p0 : -s.p0, d.y >= 0.f ? s.p1 : -s.p1, d.z >= 0.f ? s.p2 : -s.p2);
    case PBP_SHAPE_CYLINDER: {
        const float n2 = fmaf(d.x, d.x, d.y * d.y);
        const float k = n2 > 0.f ? s.p0 * rsqrtf(n2) : 0.f;
        return make_float3(d.x * k, d.y * k, d.z >= 0.f ? s.p1 : -s.p1);
    }
    default:
        return s.cell_off ? support_convex_map(s, d) : support_convex(s.verts, s.nverts, d);
    }
}

// ---------------------------------------------------------------- simplex sub-algorithms
// Each returns the point of the sub-simplex closest to the origin in `v` and a bit mask of the
// vertices that support it (bit i = i-th argument).

__device__ __forceinline__ int closest_on_segment(float3 A, float3 B, float3 &v) {
    const float3 ab = sub3(B, A);
    const float den = dot3(ab, ab);
    const float t = den > 0.f ? -dot3(A, ab) / den : 0.f;
    if (t <= 0.f) { v = A; return 1; }
    if (t >= 1.f) { v = B; return 2; }
    v = madd3(A, ab, t);
    return 3;
}

// Deliberately NOT inlined: it is called from several places of the simplex phase, and one copy
// keeps the narrowphase kernel inside the instruction cache.
__device__ __forceinline__ int closest_on_triangle_impl(float3 A, float3 B, float3 C, float3 &v) {
    const float3 ab = sub3(B, A), ac = sub3(C, A);
    const float3 n = cross3(ab, ac);
    const float nn = dot3(n, n);
    // Sliver / collinear triangle (sin^2 of the apex angle below 1e-9): the Voronoi-region signs
    // below are rounding noise in FP32, so use the best of the three edges instead.  This
    // arises with curved supports (near-duplicate points on an arc) and in planar models.
    if (!(nn > 1e-9f * dot3(ab, ab) * dot3(ac, ac))) {
        float3 v1, v2, v3;
        const int m1 = closest_on_segment(A, B, v1);
        const int m2 = closest_on_segment(A, C, v2);
        const int m3 = closest_on_segment(B, C, v3);
        const float e1 = dot3(v1, v1), e2 = dot3(v2, v2), e3 = dot3(v3, v3);
        if (e1 <= e2 && e1 <= e3) { v = v1; return m1; }
        if (e2 <= e3) { v = v2; return (m2 & 1) | ((m2 & 2) << 1); }
        v = v3;
        return m3 << 1;
    }
    const float d1 = -dot3(ab, A), d2 = -dot3(ac, A);
    if (d1 <= 0.f && d2 <= 0.f) { v = A; return 1; }
    const float d3 = -dot3(ab, B), d4 = -dot3(ac, B);
    if (d3 >= 0.f && d4 <= d3) { v = B; return 2; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float den = d1 - d3;
        v = madd3(A, ab, den > 0.f ? d1 / den : 0.f);
        return 3;
    }
    const float d5 = -dot3(ab, C), d6 = -dot3(ac, C);
    if (d6 >= 0.f && d5 <= d6) { v = C; return 4; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        const float den = d2 - d6;
        v = madd3(A, ac, den > 0.f ? d2 / den : 0.f);
        return 5;
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float den = (d4 - d3) + (d5 - d6);
        v = madd3(B, sub3(C, B), den > 0.f ? (d4 - d3) / den : 0.f);
        return 6;
    }
    // interior: project the origin on the supporting plane (better conditioned in FP32 than
    // recombining the vertices with barycentric weights)
    const float k = dot3(A, n) / nn;
    v = make_float3(n.x * k, n.y * k, n.z * k);
    return 7;
}

// Out-of-line entry point (one copy in the kernel): closest point in xyz, vertex mask in w.
__device__ __noinline__ float4 closest_on_triangle_packed(float3 A, float3 B, float3 C) {
    float3 v;
    const int m = closest_on_triangle_impl(A, B, C, v);
    return make_float4(v.x, v.y, v.z, __int_as_float(m));
}
__device__ __forceinline__ int closest_on_triangle(float3 A, float3 B, float3 C, float3 &v) {
    const float4 r = closest_on_triangle_packed(A, B, C);
    v = make_float3(r.x, r.y, r.z);
    return __float_as_int(r.w);
}

// Tetrahedron (A,B,C,D).  Returns 0 when the origin is inside (or on) it.
__device__ __forceinline__ int closest_on_tetrahedron(float3 A, float3 B, float3 C, float3 D, float3 &v) {
    float best = FLT_MAX;
    int bmask = 0;
    float3 bv = make_float3(0.f, 0.f, 0.f);
    bool any_outside = false;
    // face f omits vertex f; remap[] lifts the 3-bit triangle mask to the 4-bit tetra mask
#pragma unroll
    for (int f = 0; f < 4; f++) {
        const float3 P = f == 0 ? B : A;
        const float3 Q = f <= 1 ? C : B;
        const float3 R = f == 3 ? C : D;
        const float3 O = f == 0 ? A : (f == 1 ? B : (f == 2 ? C : D));
        const float3 n = cross3(sub3(Q, P), sub3(R, P));
        const float3 op = sub3(O, P);
        const float sd = dot3(n, op);          // side of the omitted vertex
        const float so = -dot3(n, P);          // side of the origin
        // a (nearly) flat tetrahedron has no reliable inside: its faces are all candidates
        const bool flat = sd * sd <= 1e-10f * dot3(n, n) * dot3(op, op);
        if (!flat && so * sd > 0.f) continue;  // origin strictly on the inner side of this face
        any_outside = true;
        float3 tv;
        const int m = closest_on_triangle(P, Q, R, tv);
        const float d2 = dot3(tv, tv);
        if (d2 < best) {
            best = d2;
            bv = tv;
            // triangle vertices (P,Q,R) as tetra indices
            const int ip = f == 0 ? 1 : 0, iq = f <= 1 ? 2 : 1, ir = f == 3 ? 2 : 3;
            bmask = ((m & 1) << ip) | (((m >> 1) & 1) << iq) | (((m >> 2) & 1) << ir);
        }
    }
    if (!any_outside) return 0;
    v = bv;
    return bmask;
}

struct GjkOut {
    float dist;   // distance between the cores (0 when they overlap); in boolean mode only
                  // meaningful as "<= margin" / "> margin"
    int hit;      // dist <= margin
    int iters;
    int far;      // finished through the separating-plane bound: dist is a LOWER bound beyond the cut
    int enclosed; // the tetrahedron (W0, W1, W2, w) contains the origin: the cores overlap
};

#define PBP_GJK_MAX_ITERS 48
#define PBP_GJK_TOL 1e-6f       // ub - lb convergence (metres)
#define PBP_GJK_EPS_ABS2 2.5e-13f  // |v|^2 below this counts as contact (|v| < 5e-7 m)

// Per-item GJK state: lives in registers of the owning lane between phases, so a warp can run
// "the support phase for every lane that needs one" / "the simplex phase for every lane that
// needs one" and refill finished lanes with new items.
struct GjkState {
    float3 v;            // current closest point of the simplex (search direction is -v)
    float vv;            // |v|^2
    float3 W0, W1, W2;   // simplex vertices (Minkowski-difference points), n of them valid
    float3 w;            // newest support point, handed from the support phase to the simplex phase
    int n, it, stalls;
    // TRACK only (witness points): the support points of shape A behind W0..W2 and behind w, in
    // A's frame.  Kernels that never ask for witness points never touch these (dead registers).
    float3 A0, A1, A2, a;
};

__device__ __forceinline__ void gjk_init(GjkState &s, float3 v0) {
    s.v = v0;
    s.vv = dot3(v0, v0);
    if (!(s.vv > 1e-12f)) { s.v = make_float3(1.f, 0.f, 0.f); s.vv = 1.f; }
    s.W0 = s.W1 = s.W2 = s.w = make_float3(0.f, 0.f, 0.f);
    s.A0 = s.A1 = s.A2 = s.a = make_float3(0.f, 0.f, 0.f);
    s.n = 0; s.it = 0; s.stalls = 0;
}

// Initial search direction v0 (a point of A minus a point of B, in A's frame; T = placement of B in
// A's frame).  Measured on config 2 (host emulation of this arithmetic on the undecided items):
//  * against a BOX, the point of the box closest to the other shape's centre instead of the centre
//    line: a flat box (the ground plate) is badly served by the centre line, the closest-point
//    direction is the box's own separating axis -- the first iteration then separates 55 % instead
//    of 15 % of the undecided box items;
//  * for elongated hulls, the closest points of the two shapes' long axes (the segments of their
//    enclosing capsules) instead of the proxy centres: 92 % instead of 64 % of the undecided
//    link5 - finger items, the narrowphase's most frequent pairs.
__device__ __forceinline__ float3 gjk_start_direction(const DevGeom &GA, const DevGeom &GB, const float *T) {
    const float3 ca = make_float3(GA.centre[0], GA.centre[1], GA.centre[2]);
    const float3 cb = se3_act(T, GB.centre[0], GB.centre[1], GB.centre[2]);
    float3 v = sub3(ca, cb);
    if (GA.shape == PBP_SHAPE_BOX) {
        const float3 cl = make_float3(fminf(fmaxf(cb.x, -GA.par[0]), GA.par[0]), fminf(fmaxf(cb.y, -GA.par[1]), GA.par[1]),
                                      fminf(fmaxf(cb.z, -GA.par[2]), GA.par[2]));
        const float3 d = sub3(cl, cb);
        if (dot3(d, d) > 1e-12f) v = d;            // centre of B outside the box
    } else if (GB.shape == PBP_SHAPE_BOX) {
        const float3 l = se3_act_inv(T, ca);       // centre of A in the box frame
        const float3 cl = make_float3(fminf(fmaxf(l.x, -GB.par[0]), GB.par[0]), fminf(fmaxf(l.y, -GB.par[1]), GB.par[1]),
                                      fminf(fmaxf(l.z, -GB.par[2]), GB.par[2]));
        const float3 d = sub3(ca, se3_act(T, cl.x, cl.y, cl.z));
        if (dot3(d, d) > 1e-12f) v = d;
    } else if (GA.has_axis || GB.has_axis) {
        // closest points of segment [p, p + d1] (A's axis, or its centre) and [q, q + d2] (B's)
        const float3 p = GA.has_axis ? make_float3(GA.axis0[0], GA.axis0[1], GA.axis0[2]) : ca;
        const float3 d1 = GA.has_axis ? make_float3(GA.axis1[0] - GA.axis0[0], GA.axis1[1] - GA.axis0[1], GA.axis1[2] - GA.axis0[2])
                                      : make_float3(0.f, 0.f, 0.f);
        const float3 q = GB.has_axis ? se3_act(T, GB.axis0[0], GB.axis0[1], GB.axis0[2]) : cb;
        const float3 q1 = GB.has_axis ? se3_act(T, GB.axis1[0], GB.axis1[1], GB.axis1[2]) : cb;
        const float3 d2 = sub3(q1, q);
        const float3 r = sub3(p, q);
        const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), c = dot3(d1, r), b = dot3(d1, d2);
        float s = 0.f, t = 0.f;
        if (a > 0.f && e > 0.f) {
            const float den = a * e - b * b;
            s = den > 1e-12f * a * e ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
            t = (b * s + f) / e;
            if (t < 0.f) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
            else if (t > 1.f) { t = 1.f; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
        } else if (a > 0.f) {
            s = fminf(fmaxf(-c / a, 0.f), 1.f);
        } else if (e > 0.f) {
            t = fminf(fmaxf(f / e, 0.f), 1.f);
        }
        const float3 d = sub3(madd3(p, d1, s), madd3(q, d2, t));
        if (dot3(d, d) > 1e-12f) v = d;
    }
    return v;
}

// Support phase of one GJK iteration: new support point of A - B in direction -v, then the
// termination tests that only need it.  T: placement of B in A's frame.  Returns true when the
// item is finished (out is then valid); otherwise s.w holds the point for the simplex phase.
// WANT_DIST = false: finish as soon as "distance > cut" or "distance <= margin" is certain, where
//                    cut = max(bound, margin): solids pass bound = 0 (cut = margin); surface geometries
//                    run on their inner cores and pass margin + eps, eps = how far the hulls may lie
//                    outside the cores -- an item that converges between margin and cut is UNDECIDED
//                    (out.hit = 0, out.far = 0) and goes to the triangle stage (pbp_surface.cuh).
// WANT_DIST = true : run to convergence unless the pair is certainly farther than `bound`.
// EARLY_UNDECIDED (boolean queries on cores, bound > margin): also finish -- UNDECIDED, out.hit = out.far = 0 -- as soon
//                    as the cores are certainly farther apart than the margin without being beyond the cut.  Nobody
//                    needs the converged core distance of such an item (the hulls and, if need be, the triangles
//                    decide it, pbp_surface.cuh), and converging on it in FP32 took up to 48 iterations: the lanes doing
//                    that were the tail of the narrowphase launch.
template <bool WANT_DIST, bool TRACK = false, bool EARLY_UNDECIDED = false>
__device__ __forceinline__ bool gjk_support_phase(const ShapeRef &A, const ShapeRef &B, const float *T, GjkState &s, float margin,
                                                  float bound, GjkOut &out) {
    const float3 v = s.v;
    const float vv = s.vv;
    const int n = s.n;
    const float cut = WANT_DIST ? bound : fmaxf(bound, margin);
    const float3 db = make_float3(fmaf(T[0], v.x, fmaf(T[3], v.y, T[6] * v.z)), fmaf(T[1], v.x, fmaf(T[4], v.y, T[7] * v.z)),
                                  fmaf(T[2], v.x, fmaf(T[5], v.y, T[8] * v.z)));
    // one copy of the support code serves both shapes (instruction-cache footprint)
    float3 sa = make_float3(0.f, 0.f, 0.f), sl = sa;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        ShapeRef S;
        S.shape = side ? B.shape : A.shape;
        S.p0 = side ? B.p0 : A.p0; S.p1 = side ? B.p1 : A.p1; S.p2 = side ? B.p2 : A.p2;
        S.verts = side ? B.verts : A.verts; S.nverts = side ? B.nverts : A.nverts;
        S.cell_off = side ? B.cell_off : A.cell_off; S.cell_list = side ? B.cell_list : A.cell_list;
        const float3 r = support_local(S, side ? db : make_float3(-v.x, -v.y, -v.z));
        if (side) sl = r; else sa = r;
    }
    const float3 sb = se3_act(T, sl.x, sl.y, sl.z);
    const float3 w = sub3(sa, sb);
    const float vw = dot3(v, w);
#ifdef PBP_GJK_DEBUG
    printf("it %d n %d v (%g %g %g) |v| %.9g w (%g %g %g) vw/|v| %.9g\n", s.it, n, v.x, v.y, v.z, sqrtf(vv), w.x, w.y, w.z, vw / sqrtf(vv));
#endif
    out.iters = ++s.it;
    out.far = 0;
    out.enclosed = 0;
    // <v,w>/|v| is a lower bound of the distance for ANY direction v
    if (vw > 0.f && vw * vw > cut * cut * vv) {
        out.dist = vw * rsqrtf(vv);
        out.hit = 0;
        out.far = 1;
        return true;
    }
    if (EARLY_UNDECIDED && !WANT_DIST && vw > 0.f && vw * vw > margin * margin * vv * (1.f + 1e-5f)) {
        out.dist = vw * rsqrtf(vv);   // a lower bound beyond the margin, not beyond the cut
        out.hit = 0;
        return true;
    }
    if (n > 0) {
        bool done = vv - vw <= PBP_GJK_TOL * sqrtf(vv);   // v is (numerically) the closest point
        // w (numerically) already in the simplex: v cannot improve any further
        const float near2 = 1e-10f * dot3(w, w);
        const float3 e0 = sub3(w, s.W0), e1 = sub3(w, s.W1), e2 = sub3(w, s.W2);
        if (dot3(e0, e0) <= near2 || (n > 1 && dot3(e1, e1) <= near2) || (n > 2 && dot3(e2, e2) <= near2)) done = true;
        if (done) {
            out.dist = sqrtf(vv);
            out.hit = out.dist <= margin;
            return true;
        }
    }
    s.w = w;
    if (TRACK) s.a = sa;
    return false;
}

// Simplex phase: append s.w, find the point of the simplex closest to the origin, reduce the
// simplex to the face that supports it.  Returns true when the item is finished.
template <bool WANT_DIST, bool TRACK = false>
__device__ __forceinline__ bool gjk_simplex_phase(GjkState &s, float margin, GjkOut &out) {
    const float3 w = s.w;
    const float vv = s.vv;
    const int n = s.n;
    out.iters = s.it;
    out.far = 0;
    out.enclosed = 0;
    bool done = false, enclosed = false;
    {
        // append w and solve for the closest point of the simplex
        float3 nv, W3 = w;
        int mask;
        if (TRACK) {
            if (n == 0) s.A0 = s.a;
            else if (n == 1) s.A1 = s.a;
            else if (n == 2) s.A2 = s.a;
        }
        if (n == 0) { s.W0 = w; nv = w; mask = 1; }
        else if (n == 1) { s.W1 = w; mask = closest_on_segment(s.W0, s.W1, nv); }
        else {
            // Triangle (n == 2) and tetrahedron (n == 3) share one code path so that the lanes of a
            // warp, which sit at different simplex sizes, stay converged: the tetrahedron visits its
            // four faces (face f omits vertex f), the triangle only "face 3" = (W0, W1, W2).
            if (n == 2) s.W2 = w;
            float best = FLT_MAX;
            int bmask = 0;
            float3 bv = make_float3(0.f, 0.f, 0.f);
            bool any_outside = false;
#pragma unroll 1
            for (int f = (n == 2 ? 3 : 0); f < 4; f++) {
                const float3 P = f == 0 ? s.W1 : s.W0;
                const float3 Q = f <= 1 ? s.W2 : s.W1;
                const float3 R = f == 3 ? s.W2 : W3;
                if (n == 3) {
                    const float3 O = f == 0 ? s.W0 : (f == 1 ? s.W1 : (f == 2 ? s.W2 : W3));
                    const float3 nf = cross3(sub3(Q, P), sub3(R, P));
                    const float3 op = sub3(O, P);
                    const float sd = dot3(nf, op);          // side of the omitted vertex
                    const float so = -dot3(nf, P);          // side of the origin
                    // a (nearly) flat tetrahedron has no reliable inside: its faces are all candidates
                    const bool flat = sd * sd <= 1e-10f * dot3(nf, nf) * dot3(op, op);
                    if (!flat && so * sd > 0.f) continue;   // origin strictly on the inner side of this face
                }
                any_outside = true;
                float3 tv;
                const int m = closest_on_triangle(P, Q, R, tv);
                const float d2 = dot3(tv, tv);
                if (d2 < best) {
                    best = d2;
                    bv = tv;
                    const int ip = f == 0 ? 1 : 0, iq = f <= 1 ? 2 : 1, ir = f == 3 ? 2 : 3;
                    bmask = ((m & 1) << ip) | (((m >> 1) & 1) << iq) | (((m >> 2) & 1) << ir);
                }
            }
            if (!any_outside) { enclosed = true; bv = make_float3(0.f, 0.f, 0.f); }
            nv = bv;
            mask = bmask;
        }
        float nvv = dot3(nv, nv);
        if (!enclosed && n >= 2 && nvv >= vv) {
            // Backup step: sliver simplices (curved supports, near-duplicate points) can defeat the
            // region tests in FP32; retry with the sub-simplices that contain the new vertex.
            if (n == 2) {
                float3 c1, c2;
                const int a1 = closest_on_segment(s.W0, s.W2, c1), a2 = closest_on_segment(s.W1, s.W2, c2);
                const float e1 = dot3(c1, c1), e2 = dot3(c2, c2);
                if (e1 < nvv && e1 <= e2) { nv = c1; mask = (a1 & 1) | ((a1 & 2) << 1); }
                else if (e2 < nvv) { nv = c2; mask = a2 << 1; }
            } else {
                // the three faces that contain the new vertex W3: (W0,W1,W3), (W0,W2,W3), (W1,W2,W3)
#pragma unroll 1
                for (int f = 0; f < 3; f++) {
                    const float3 P = f == 2 ? s.W1 : s.W0;
                    const float3 Q = f == 0 ? s.W1 : s.W2;
                    float3 c;
                    const int a = closest_on_triangle(P, Q, W3, c);
                    const float e = dot3(c, c);
                    if (e < nvv) {
                        nvv = e;
                        nv = c;
                        const int ip = f == 2 ? 1 : 0, iq = f == 0 ? 1 : 2;
                        mask = ((a & 1) << ip) | (((a >> 1) & 1) << iq) | (((a >> 2) & 1) << 3);
                    }
                }
            }
            nvv = dot3(nv, nv);
        }
        if (enclosed || nvv <= PBP_GJK_EPS_ABS2) {
            out.dist = 0.f;
            out.hit = 1;
            out.enclosed = enclosed ? 1 : 0;
            return true;
        }
        if (!WANT_DIST && nvv <= margin * margin) {  // upper bound already within margin
            out.dist = sqrtf(nvv);
            out.hit = 1;
            return true;
        }
        // |v| must decrease; a decrease below FP32 resolution (large flat faces far from the origin)
        // is still real progress -- the direction turns and the next support finds a new vertex --
        // so tolerate a few such steps, but stop on a genuinely worse point or when it persists.
        if (n > 0 && nvv >= vv) {
            if (nvv > vv * 1.00001f || ++s.stalls > 4) done = true;
        } else {
            s.stalls = 0;
        }
        if (!done) {
            s.v = nv;
            s.vv = nvv;
            // compact the kept vertices to the front, branch-free (select chains on static registers)
            const int m1 = mask & (mask - 1), m2 = m1 & (m1 - 1);
            const int i0 = __ffs(mask) - 1, i1 = __ffs(m1) - 1, i2 = __ffs(m2) - 1;
            auto pick = [&](int i) { return i == 0 ? s.W0 : (i == 1 ? s.W1 : (i == 2 ? s.W2 : W3)); };
            const float3 n0 = pick(i0), n1 = pick(i1), n2 = pick(i2);
            s.W0 = n0; s.W1 = n1; s.W2 = n2;
            if (TRACK) {
                auto pick_a = [&](int i) { return i == 0 ? s.A0 : (i == 1 ? s.A1 : (i == 2 ? s.A2 : s.a)); };
                const float3 a0 = pick_a(i0), a1 = pick_a(i1), a2 = pick_a(i2);
                s.A0 = a0; s.A1 = a1; s.A2 = a2;
            }
            s.n = __popc(mask);
            if (s.it >= PBP_GJK_MAX_ITERS) done = true;
        }
    }
    if (done) {
        out.dist = sqrtf(s.vv);
        out.hit = out.dist <= margin;
        return true;
    }
    return false;
}

// One full GJK iteration (support phase, then simplex phase).
template <bool WANT_DIST, bool TRACK = false>
__device__ __forceinline__ bool gjk_step(const ShapeRef &A, const ShapeRef &B, const float *T, GjkState &s, float margin,
                                         float bound, GjkOut &out) {
    if (gjk_support_phase<WANT_DIST, TRACK>(A, B, T, s, margin, bound, out)) return true;
    return gjk_simplex_phase<WANT_DIST, TRACK>(s, margin, out);
}

// Witness points on the cores (A's frame) of a TRACKed, finished, non-overlapping item:
// pa = sum(lambda_i A_i) and pb = sum(lambda_i (A_i - W_i)), the same convex combination of the
// support points of A and of B whose Minkowski image sum(lambda_i W_i) is the closest point s.v.
// Both are convex combinations of points OF their shapes (so they lie in them whatever the
// residual error of v); the barycentric weights are recomputed from W and v.
__device__ __forceinline__ void gjk_witness(const GjkState &s, float3 &pa, float3 &pb) {
    float l1 = 0.f, l2 = 0.f;   // weights of vertex 1 and 2 (vertex 0 gets the rest)
    if (s.n == 2) {
        const float3 e = sub3(s.W1, s.W0);
        const float den = dot3(e, e);
        l1 = den > 0.f ? dot3(sub3(s.v, s.W0), e) / den : 0.f;
        l1 = fminf(fmaxf(l1, 0.f), 1.f);
    } else if (s.n >= 3) {
        // triangle: v = sum(lam_i W_i), least squares in the plane, solved from the corner with the
        // widest angle-sine.  GJK on curved supports ends on needle triangles (two rim points of a
        // cylinder 0.4 mm apart, the third 35 cm away): from the needle's tip the Gram determinant
        // a11 a22 - a12^2 cancels to a few FP32 digits (weights wrong by 10 %, witnesses by
        // centimetres), from an end of the short edge it is well conditioned.
        const float3 Wc[3] = {s.W0, s.W1, s.W2};
        float lam[3] = {1.f, 0.f, 0.f};
        float best_rel = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float3 f1 = sub3(Wc[(c + 1) % 3], Wc[c]), f2 = sub3(Wc[(c + 2) % 3], Wc[c]), rc = sub3(s.v, Wc[c]);
            const float c11 = dot3(f1, f1), c12 = dot3(f1, f2), c22 = dot3(f2, f2);
            const float dc = c11 * c22 - c12 * c12;
            const float rel = c11 * c22 > 0.f ? dc / (c11 * c22) : 0.f;
            if (rel > best_rel) {
                best_rel = rel;
                const float g1 = dot3(rc, f1), g2 = dot3(rc, f2);
                const float u = (g1 * c22 - g2 * c12) / dc, w = (g2 * c11 - g1 * c12) / dc;
                lam[(c + 1) % 3] = u; lam[(c + 2) % 3] = w; lam[c] = 1.f - u - w;
            }
        }
        const float3 e1 = sub3(s.W1, s.W0), e2 = sub3(s.W2, s.W0), r = sub3(s.v, s.W0);
        const float a11 = dot3(e1, e1), a22 = dot3(e2, e2);
        const float b1 = dot3(r, e1), b2 = dot3(r, e2);
        bool ok = false;
        if (best_rel > 1e-6f) {
            l1 = lam[1]; l2 = lam[2];
            ok = l1 >= -1e-4f && l2 >= -1e-4f && l1 + l2 <= 1.0001f;
            l1 = fminf(fmaxf(l1, 0.f), 1.f);
            l2 = fminf(fmaxf(l2, 0.f), 1.f - l1);
        }
        if (!ok) {
            // Sliver triangle (curved supports: two rim points of a cylinder next to each other) or
            // weights that needed clamping: v lies on one of the edges to working precision -- take
            // the edge whose closest point reproduces v best, so that sum(l_i W_i) = v holds and the
            // two witnesses stay |v| apart.
            const float3 e3 = sub3(s.W2, s.W1);
            const float a33 = dot3(e3, e3);
            const float t1 = a11 > 0.f ? fminf(fmaxf(b1 / a11, 0.f), 1.f) : 0.f;                      // edge 0-1
            const float t2 = a22 > 0.f ? fminf(fmaxf(b2 / a22, 0.f), 1.f) : 0.f;                      // edge 0-2
            const float t3 = a33 > 0.f ? fminf(fmaxf(dot3(sub3(s.v, s.W1), e3) / a33, 0.f), 1.f) : 0.f; // edge 1-2
            const float3 r1 = sub3(r, make_float3(e1.x * t1, e1.y * t1, e1.z * t1));
            const float3 r2 = sub3(r, make_float3(e2.x * t2, e2.y * t2, e2.z * t2));
            const float3 r3 = sub3(sub3(s.v, s.W1), make_float3(e3.x * t3, e3.y * t3, e3.z * t3));
            const float q1 = dot3(r1, r1), q2 = dot3(r2, r2), q3 = dot3(r3, r3);
            float q0 = FLT_MAX;
            if (best_rel > 1e-6f) {
                const float3 r0 = sub3(r, make_float3(e1.x * l1 + e2.x * l2, e1.y * l1 + e2.y * l2, e1.z * l1 + e2.z * l2));
                q0 = dot3(r0, r0);
            }
            if (q1 <= q0 && q1 <= q2 && q1 <= q3) { l1 = t1; l2 = 0.f; }
            else if (q2 <= q0 && q2 <= q3) { l1 = 0.f; l2 = t2; }
            else if (q3 <= q0) { l1 = 1.f - t3; l2 = t3; }
        }
    }
    pa = s.A0;
    float3 m = s.W0;   // Minkowski image of the combination
    if (s.n >= 2) { pa = madd3(pa, sub3(s.A1, s.A0), l1); m = madd3(m, sub3(s.W1, s.W0), l1); }
    if (s.n >= 3) { pa = madd3(pa, sub3(s.A2, s.A0), l2); m = madd3(m, sub3(s.W2, s.W0), l2); }
    pb = sub3(pa, m);
}

// Whole item in one call (host build used by the CPU tests; the device narrowphase drives
// gjk_step itself so that finished lanes can be refilled).
template <bool WANT_DIST>
__device__ __forceinline__ GjkOut gjk_run(const ShapeRef &A, const ShapeRef &B, const float *T, float3 v0, float margin,
                                          float bound) {
    GjkState s;
    GjkOut out;
    gjk_init(s, v0);
    while (!gjk_step<WANT_DIST>(A, B, T, s, margin, bound, out)) {
    }
    return out;
}

}  // namespace pbp
