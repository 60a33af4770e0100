// Per-pair distances with witness points and the collision-avoidance gradient built from them.
//
// Restates, for N configurations at once,
//   * pinocchio.computeDistances as the reference reads it: distanceResults[k].min_distance,
//     getNearestPoint1/2(), normal (pointing from object 1 to object 2)
//     (/root/reference/src/pyroboplan/core/utils.py:505-516, ik/nullspace_components.py:124-131);
//   * calculate_collision_vector_and_jacobians + collision_avoidance_nullspace_component
//     (/root/reference/src/pyroboplan/core/utils.py:478-538, ik/nullspace_components.py:82-142):
//       component -= (normal * dist - dist_padding) @ (Jcoll2 - Jcoll1)   for every pair with
//       dist <= dist_padding,   Jcoll = [I, -skew(r)] * LOCAL_WORLD_ALIGNED frame Jacobian, i.e.
//     the Jacobian of the witness point's linear velocity: column k = z_k x (p - o_k) for a revolute
//     joint k on the chain of the geometry's parent joint, z_k for a prismatic one.
//
// Overlapping pairs report coal's convention too: a NEGATIVE distance (the penetration depth), the
// deepest points and the normal from shape 1 to shape 2 -- from an FP32 Expanding Polytope
// Algorithm started on the GJK tetrahedron that encloses the origin (per-thread polytope of at most
// EPA_MAXV vertices in local memory; this is not the hot path).  Scalar FP32, like the rest of the
// collision path.
#pragma once
#include "pbp_gjk.cuh"

namespace pbp {

struct WitnessOut {
    float dist;      // signed surface distance: > 0 separated, < 0 penetration depth; lower bound when `far`
    int hit;         // shapes overlap (dist <= 0)
    int far;         // certainly farther than the cutoff: no witness points
    float3 pa, pb;   // witness points on the surfaces, in A's frame
    float3 n;        // unit normal from A to B, in A's frame (0 when hit / far)
};

constexpr int EPA_MAXV = 40;
constexpr int EPA_MAXF = 2 * EPA_MAXV;     // a triangulated polytope with V vertices has 2V - 4 faces
constexpr int EPA_MAXIT = EPA_MAXV - 4;

struct EpaFace {
    float nx, ny, nz, d;   // outward unit normal and plane distance from the origin (d >= 0)
    unsigned char v0, v1, v2, alive;
};

__device__ __forceinline__ bool epa_make_face(EpaFace &f, int i0, int i1, int i2, const float3 *W) {
    const float3 n = cross3(sub3(W[i1], W[i0]), sub3(W[i2], W[i0]));
    const float len2 = dot3(n, n);
    f.v0 = (unsigned char)i0; f.v1 = (unsigned char)i1; f.v2 = (unsigned char)i2; f.alive = 1;
    if (!(len2 > 1e-30f)) { f.nx = f.ny = f.nz = 0.f; f.d = FLT_MAX; return false; }
    const float inv = rsqrtf(len2);
    f.nx = n.x * inv; f.ny = n.y * inv; f.nz = n.z * inv;
    f.d = f.nx * W[i0].x + f.ny * W[i0].y + f.nz * W[i0].z;
    if (f.d < 0.f) {   // orient away from the origin, which is inside
        f.v1 = (unsigned char)i2; f.v2 = (unsigned char)i1;
        f.nx = -f.nx; f.ny = -f.ny; f.nz = -f.nz; f.d = -f.d;
    }
    return true;
}

// Penetration of the CORES, started from the GJK simplex that contains the origin: the enclosing
// tetrahedron (s.W0, s.W1, s.W2, s.w), or -- when GJK stopped on a lower-dimensional simplex through
// the origin, as it does in planar models -- a bipyramid built around that triangle / segment with
// extra support points.  Returns false when no polytope with the origin inside can be started.
__device__ __noinline__ bool epa_run(const ShapeRef &A, const ShapeRef &B, const float *T, const GjkState &s, bool enclosed, float &depth,
                                     float3 &n, float3 &pa, float3 &pb) {
    float3 W[EPA_MAXV], Av[EPA_MAXV];
    EpaFace F[EPA_MAXF];
    unsigned char eu[3 * EPA_MAXF], ev[3 * EPA_MAXF];
    auto support_pair = [&](float3 d, float3 &w, float3 &sa) {
        const float3 db = make_float3(-(T[0] * d.x + T[3] * d.y + T[6] * d.z), -(T[1] * d.x + T[4] * d.y + T[7] * d.z),
                                      -(T[2] * d.x + T[5] * d.y + T[8] * d.z));
        sa = support_local(A, d);
        const float3 sl = support_local(B, db);
        w = sub3(sa, se3_act(T, sl.x, sl.y, sl.z));
    };
    int nv, nf;
    const int npts = enclosed ? 4 : s.n + 1;     // the simplex phase appended s.w to the s.n points it held
    if (npts >= 4) {
        W[0] = s.W0; W[1] = s.W1; W[2] = s.W2; W[3] = s.w;
        Av[0] = s.A0; Av[1] = s.A1; Av[2] = s.A2; Av[3] = s.a;
        nv = 4; nf = 4;
        epa_make_face(F[0], 0, 1, 2, W);
        epa_make_face(F[1], 0, 1, 3, W);
        epa_make_face(F[2], 0, 2, 3, W);
        epa_make_face(F[3], 1, 2, 3, W);
    } else if (npts == 3) {
        // triangle through the origin: apexes along +- its normal
        W[0] = s.W0; W[1] = s.W1; W[2] = s.W2; Av[0] = s.A0; Av[1] = s.A1; Av[2] = s.A2;   // W2 / A2 hold the appended point
        const float3 nn = cross3(sub3(W[1], W[0]), sub3(W[2], W[0]));
        if (!(dot3(nn, nn) > 1e-30f)) return false;
        support_pair(nn, W[3], Av[3]);
        support_pair(make_float3(-nn.x, -nn.y, -nn.z), W[4], Av[4]);
        nv = 5; nf = 6;
        epa_make_face(F[0], 0, 1, 3, W); epa_make_face(F[1], 1, 2, 3, W); epa_make_face(F[2], 2, 0, 3, W);
        epa_make_face(F[3], 0, 1, 4, W); epa_make_face(F[4], 1, 2, 4, W); epa_make_face(F[5], 2, 0, 4, W);
    } else if (npts == 2) {
        // segment through the origin: a ring of three support points around it
        W[0] = s.W0; W[1] = s.W1; Av[0] = s.A0; Av[1] = s.A1;                              // W1 / A1 hold the appended point
        const float3 u = sub3(W[1], W[0]);
        if (!(dot3(u, u) > 1e-30f)) return false;
        const float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
        const float3 e = (ax <= ay && ax <= az) ? make_float3(1.f, 0.f, 0.f) : (ay <= az ? make_float3(0.f, 1.f, 0.f) : make_float3(0.f, 0.f, 1.f));
        const float3 d1 = cross3(u, e), d2 = cross3(u, d1);
        const float k1 = rsqrtf(dot3(d1, d1)), k2 = rsqrtf(dot3(d2, d2));
        const float3 a1 = make_float3(d1.x * k1, d1.y * k1, d1.z * k1), a2 = make_float3(d2.x * k2, d2.y * k2, d2.z * k2);
        support_pair(a1, W[2], Av[2]);
        support_pair(make_float3(-0.5f * a1.x + 0.8660254f * a2.x, -0.5f * a1.y + 0.8660254f * a2.y, -0.5f * a1.z + 0.8660254f * a2.z), W[3], Av[3]);
        support_pair(make_float3(-0.5f * a1.x - 0.8660254f * a2.x, -0.5f * a1.y - 0.8660254f * a2.y, -0.5f * a1.z - 0.8660254f * a2.z), W[4], Av[4]);
        nv = 5; nf = 6;
        epa_make_face(F[0], 2, 3, 0, W); epa_make_face(F[1], 3, 4, 0, W); epa_make_face(F[2], 4, 2, 0, W);
        epa_make_face(F[3], 2, 3, 1, W); epa_make_face(F[4], 3, 4, 1, W); epa_make_face(F[5], 4, 2, 1, W);
    } else {
        return false;
    }
    int best = -1;
    for (int iter = 0; iter <= EPA_MAXIT; iter++) {
        best = -1;
        float bd = FLT_MAX;
        for (int f = 0; f < nf; f++)
            if (F[f].alive && F[f].d < bd) { bd = F[f].d; best = f; }
        if (best < 0) return false;
        const float3 d = make_float3(F[best].nx, F[best].ny, F[best].nz);
        float3 w, sa;
        support_pair(d, w, sa);   // support of A - B along d
        const float proj = dot3(w, d);
        if (proj - bd <= 1e-6f * (1.f + bd) || nv >= EPA_MAXV || iter == EPA_MAXIT) break;
        W[nv] = w; Av[nv] = sa;
        int ne = 0;
        for (int f = 0; f < nf; f++) {
            if (!F[f].alive) continue;
            const float3 rel = sub3(w, W[F[f].v0]);
            const bool sees = F[f].d == FLT_MAX || (F[f].nx * rel.x + F[f].ny * rel.y + F[f].nz * rel.z) > 1e-7f;
            if (sees) {
                F[f].alive = 0;
                eu[ne] = F[f].v0; ev[ne] = F[f].v1; ne++;
                eu[ne] = F[f].v1; ev[ne] = F[f].v2; ne++;
                eu[ne] = F[f].v2; ev[ne] = F[f].v0; ne++;
            }
        }
        if (ne == 0) break;   // numerically on the surface
        bool overflow = false;
        for (int e = 0; e < ne; e++) {
            bool shared = false;
            for (int g = 0; g < ne; g++)
                if (g != e && eu[g] == ev[e] && ev[g] == eu[e]) { shared = true; break; }
            if (shared) continue;
            int slot = -1;
            for (int f = 0; f < nf; f++)
                if (!F[f].alive) { slot = f; break; }
            if (slot < 0) {
                if (nf >= EPA_MAXF) { overflow = true; break; }
                slot = nf++;
            }
            epa_make_face(F[slot], eu[e], ev[e], nv, W);
        }
        nv++;
        if (overflow) break;
    }
    if (best < 0 || F[best].d == FLT_MAX) return false;
    // among the (coplanar) faces at the minimal distance take the one that contains the projection of the origin
    const float dmin = F[best].d;
    float bu = 0.f, bw = 0.f, bviol = FLT_MAX;
    int bf = best;
    for (int f = 0; f < nf; f++) {
        if (!F[f].alive || !(F[f].d <= dmin + 1e-6f * (1.f + dmin))) continue;
        const float3 m = make_float3(F[f].nx * F[f].d, F[f].ny * F[f].d, F[f].nz * F[f].d);
        const float3 e1 = sub3(W[F[f].v1], W[F[f].v0]), e2 = sub3(W[F[f].v2], W[F[f].v0]), r = sub3(m, W[F[f].v0]);
        const float a11 = dot3(e1, e1), a12 = dot3(e1, e2), a22 = dot3(e2, e2), b1 = dot3(r, e1), b2 = dot3(r, e2);
        const float det = a11 * a22 - a12 * a12;
        if (!(det > 0.f)) continue;
        const float u = (b1 * a22 - b2 * a12) / det, v = (b2 * a11 - b1 * a12) / det;
        const float viol = fmaxf(fmaxf(-u, -v), u + v - 1.f);
        if (viol < bviol) { bviol = viol; bu = u; bw = v; bf = f; }
    }
    bu = fmaxf(bu, 0.f); bw = fmaxf(bw, 0.f);
    if (bu + bw > 1.f) { const float t = bu + bw; bu /= t; bw /= t; }
    const EpaFace &f = F[bf];
    depth = f.d;
    n = make_float3(f.nx, f.ny, f.nz);
    pa = madd3(madd3(Av[f.v0], sub3(Av[f.v1], Av[f.v0]), bu), sub3(Av[f.v2], Av[f.v0]), bw);
    // the witness on B is the one on A moved by the penetration vector n * depth.  (The barycentric
    // point of the face equals n * depth on a convex polytope; on curved supports the 40-vertex
    // polytope has sliver faces, the projection of the origin can fall outside the closest one and
    // the clamped weights would break p2 - p1 = n * distance, which the callers rely on.)
    pb = make_float3(pa.x - n.x * depth, pa.y - n.y * depth, pa.z - n.z * depth);
    return true;
}

// Full GJK with witness tracking for one item.  radii = core inflation of A + B.
__device__ __forceinline__ WitnessOut gjk_run_witness(const ShapeRef &A, const ShapeRef &B, const float *T, float3 v0, float ra,
                                                      float rb, float cutoff) {
    GjkState s;
    GjkOut out;
    gjk_init(s, v0);
    const float radii = ra + rb;
    const float bound = cutoff + radii;
    for (;;) {
        if (gjk_support_phase<true, true>(A, B, T, s, radii, bound, out)) break;
        if (gjk_simplex_phase<true, true>(s, radii, out)) break;
    }
    WitnessOut w;
    w.pa = w.pb = w.n = make_float3(0.f, 0.f, 0.f);
    w.hit = out.hit;
    w.far = 0;
    if (out.far) {   // separating-plane bound beyond the cutoff: no converged simplex, no witness
        w.dist = out.dist - radii;
        w.far = 1;
        return w;
    }
    if (out.hit && out.dist == 0.f) {
        // the cores overlap: penetration depth of the cores by EPA, the radii add to it
        float depth = 0.f;
        float3 n = make_float3(0.f, 0.f, 0.f), pa = n, pb = n;
        if (epa_run(A, B, T, s, out.enclosed != 0, depth, n, pa, pb)) {
            w.dist = -(depth + radii);
            w.n = n;
            w.pa = madd3(pa, n, ra);
            w.pb = madd3(pb, n, -rb);
        } else {
            // touching cores (|v| below resolution): depth of the inflation only, along the last direction
            const float len = sqrtf(s.vv);
            const float inv = len > 0.f ? 1.0f / len : 0.f;
            w.n = make_float3(-s.v.x * inv, -s.v.y * inv, -s.v.z * inv);
            float3 qa, qb;
            gjk_witness(s, qa, qb);
            w.dist = -radii;
            w.pa = madd3(qa, w.n, ra);
            w.pb = madd3(qb, w.n, -rb);
        }
        return w;
    }
    const float len = sqrtf(s.vv);
    w.dist = len - radii;           // negative when only the inflated shapes overlap (cores apart)
    float3 pa, pb;
    gjk_witness(s, pa, pb);
    // normal = direction witness A -> witness B (self-consistent with the points); -v/|v| when the
    // cores (nearly) touch and that difference is rounding noise
    const float3 ab = sub3(pb, pa);
    const float lab = sqrtf(dot3(ab, ab));
    if (lab > 1e-5f) {
        const float inv = 1.0f / lab;
        w.n = make_float3(ab.x * inv, ab.y * inv, ab.z * inv);
    } else {
        const float inv = len > 0.f ? 1.0f / len : 0.f;
        w.n = make_float3(-s.v.x * inv, -s.v.y * inv, -s.v.z * inv);
    }
    w.pa = madd3(pa, w.n, ra);
    w.pb = madd3(pb, w.n, -rb);
    return w;
}

#ifdef __CUDACC__   // the kernels below are device only; the host build (tests/host_emu) stops here
__device__ __forceinline__ ShapeRef shape_ref(const DevModel &M, const DevGeom &G) {
    ShapeRef S;
    S.shape = G.shape; S.p0 = G.par[0]; S.p1 = G.par[1]; S.p2 = G.par[2];
    S.verts = M.verts + G.vert_off; S.nverts = G.vert_cnt;
    S.cell_off = G.map_off >= 0 ? M.sup_cells + G.map_off : nullptr;
    S.cell_list = M.sup_lists + G.list_off;
    return S;
}

// One thread per (pair, state) item, pair-major so the lanes of a warp share the pair's shapes.
// oMg: [N][ngeoms][12] world placements (k_fk_output).  Outputs are indexed by the CALLER's pair
// index k = pair_orig[p]:  dist[k*N + i] (signed), p1/p2/normal[(k*N + i)*3 ..] in the world frame.
__global__ void __launch_bounds__(128)
k_pair_distances(DevModel M, const float *__restrict__ oMg, int64_t n, float cutoff, float *__restrict__ dist,
                 float *__restrict__ p1, float *__restrict__ p2, float *__restrict__ normal) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * M.npairs) return;
    const int p = (int)(t / n);
    const int64_t i = t - (int64_t)p * n;
    const BpPair pr = M.bppairs[p];
    const DevGeom &GA = M.geoms[pr.a];
    const DevGeom &GB = M.geoms[pr.b];
    float Ta[12], Tb[12], T[12];
#pragma unroll
    for (int k = 0; k < 12; k++) { Ta[k] = oMg[(i * M.ngeoms + pr.a) * 12 + k]; Tb[k] = oMg[(i * M.ngeoms + pr.b) * 12 + k]; }
    se3_inv_mul(Ta, Tb, T);
    const float3 v0 = gjk_start_direction(GA, GB, T);
    const WitnessOut w = gjk_run_witness(shape_ref(M, GA), shape_ref(M, GB), T, v0, GA.core_radius, GB.core_radius, cutoff);
    const int64_t o = (int64_t)M.pair_orig[p] * n + i;
    dist[o] = w.dist;
    const float3 wa = se3_act(Ta, w.pa.x, w.pa.y, w.pa.z), wb = se3_act(Ta, w.pb.x, w.pb.y, w.pb.z);
    const bool valid = !w.far;
    const float3 nw = make_float3(fmaf(Ta[0], w.n.x, fmaf(Ta[1], w.n.y, Ta[2] * w.n.z)), fmaf(Ta[3], w.n.x, fmaf(Ta[4], w.n.y, Ta[5] * w.n.z)),
                                  fmaf(Ta[6], w.n.x, fmaf(Ta[7], w.n.y, Ta[8] * w.n.z)));
    if (p1) { p1[o * 3] = valid ? wa.x : 0.f; p1[o * 3 + 1] = valid ? wa.y : 0.f; p1[o * 3 + 2] = valid ? wa.z : 0.f; }
    if (p2) { p2[o * 3] = valid ? wb.x : 0.f; p2[o * 3 + 1] = valid ? wb.y : 0.f; p2[o * 3 + 2] = valid ? wb.z : 0.f; }
    if (normal) { normal[o * 3] = valid ? nw.x : 0.f; normal[o * 3 + 1] = valid ? nw.y : 0.f; normal[o * 3 + 2] = valid ? nw.z : 0.f; }
}

// component[i][:] = gain * sum over pairs with dist <= dist_padding of
//     -(normal * dist - dist_padding) @ (Jcoll2 - Jcoll1)            (thread per state)
// oMi: [N][njoints][12] world joint placements; pair arrays as written by k_pair_distances
// (caller's pair order); pairs: [npairs][2] geometry ids in the caller's order.
__global__ void __launch_bounds__(128)
k_collision_nullspace(DevModel M, const int32_t *__restrict__ pairs, const float *__restrict__ oMi, const float *__restrict__ dist,
                      const float *__restrict__ p1, const float *__restrict__ p2, const float *__restrict__ normal, int64_t n,
                      float dist_padding, float gain, float *__restrict__ component) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float comp[PBP_MAX_NQ];
    for (int k = 0; k < M.nq; k++) comp[k] = 0.f;
    for (int p = 0; p < M.npairs; p++) {
        const int64_t o = (int64_t)p * n + i;
        const float d = dist[o];
        if (!(d <= dist_padding)) continue;
        // distance_vec - dist_padding: the reference subtracts the scalar from every component
        const float vx = normal[o * 3] * d - dist_padding, vy = normal[o * 3 + 1] * d - dist_padding,
                    vz = normal[o * 3 + 2] * d - dist_padding;
#pragma unroll 1
        for (int side = 0; side < 2; side++) {
            const float *pw = (side ? p2 : p1) + o * 3;
            const float sign = side ? -1.f : 1.f;   // comp -= v @ (J2 - J1)  ->  -J2 term, +J1 term
            for (int j = M.geoms[pairs[2 * p + side]].parent; j > 0; j = M.joints[j].parent) {
                const DevJoint &J = M.joints[j];
                const float *X = oMi + (i * M.njoints + j) * 12;
                const float zx = fmaf(X[0], J.axis[0], fmaf(X[1], J.axis[1], X[2] * J.axis[2]));
                const float zy = fmaf(X[3], J.axis[0], fmaf(X[4], J.axis[1], X[5] * J.axis[2]));
                const float zz = fmaf(X[6], J.axis[0], fmaf(X[7], J.axis[1], X[8] * J.axis[2]));
                float cx = zx, cy = zy, cz = zz;
                if (J.type == PBP_JOINT_REVOLUTE) {
                    const float rx = pw[0] - X[9], ry = pw[1] - X[10], rz = pw[2] - X[11];
                    cx = zy * rz - zz * ry; cy = zz * rx - zx * rz; cz = zx * ry - zy * rx;
                }
                comp[J.idx_q] += sign * (vx * cx + vy * cy + vz * cz);
            }
        }
    }
    for (int k = 0; k < M.nq; k++) component[i * M.nq + k] = gain * comp[k];
}

#endif  // __CUDACC__

}  // namespace pbp
