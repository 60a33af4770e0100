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
// Penetration depth (EPA) is not implemented: overlapping pairs report distance 0 and a zero
// normal (the reference's coal reports a negative distance there).  Scalar FP32, like the rest
// of the collision path.
#pragma once
#include "pbp_gjk.cuh"

namespace pbp {

struct WitnessOut {
    float dist;      // surface distance (>= 0); lower bound when `far`
    int hit;         // shapes overlap (dist == 0)
    int far;         // certainly farther than the cutoff: no witness points
    float3 pa, pb;   // witness points on the surfaces, in A's frame
    float3 n;        // unit normal from A to B, in A's frame (0 when hit / far)
};

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
    if (out.hit) { w.dist = 0.f; return w; }
    if (out.far) {   // separating-plane bound beyond the cutoff: no converged simplex, no witness
        w.dist = out.dist - radii;
        w.far = 1;
        return w;
    }
    const float len = sqrtf(s.vv);
    w.dist = fmaxf(len - radii, 0.f);
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
// index k = pair_orig[p]:  dist[k*N + i], p1/p2/normal[(k*N + i)*3 ..] in the world frame.
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
    const bool valid = !w.hit && !w.far;
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
