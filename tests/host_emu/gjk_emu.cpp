// Host build of the engine's FP32 GJK (pyroboplan_b200/csrc/pbp_gjk.cuh) for CPU-side tests.
//
// TEST TOOLING ONLY: compiled by tests/ with g++ (no nvcc, no GPU), never loaded by the
// product.  It lets `pytest -m "not gpu"` exercise the exact float32 arithmetic of the device
// narrowphase (fmaf / IEEE sqrt / division are identical on host and device; only rsqrtf is
// emulated) against the float64 oracle on many random pairs.
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#include <algorithm>
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int __popc(int x) { return __builtin_popcount((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline float __int_as_float(int i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; __builtin_memcpy(&i, &f, 4); return i; }
using std::max;
using std::min;

#include "../../pyroboplan_b200/csrc/pbp_gjk.cuh"
#include "../../pyroboplan_b200/csrc/pbp_supmap.h"
#include "../../pyroboplan_b200/csrc/pbp_witness.cuh"

static_assert(pbp::SUPMAP_KD == pbp::SUPMAP_K, "support map resolution mismatch");

extern "C" {

// batch over n items sharing the shapes; T [n][12], v0 [n][3].
// use_maps != 0: convex supports go through the exact support maps (as on the device).
void emu_gjk_batch(int shapeA, const float *parA, const float *vertsA, int nA, int shapeB, const float *parB,
                   const float *vertsB, int nB, const float *T, const float *v0, int64_t n, float margin, float bound,
                   int want_dist, int use_maps, int32_t *hit, float *dist, int32_t *iters) {
    pbp::ShapeRef A, B;
    A.shape = shapeA; A.p0 = parA[0]; A.p1 = parA[1]; A.p2 = parA[2];
    A.verts = reinterpret_cast<const float4 *>(vertsA); A.nverts = nA; A.cell_off = nullptr; A.cell_list = nullptr;
    B.shape = shapeB; B.p0 = parB[0]; B.p1 = parB[1]; B.p2 = parB[2];
    B.verts = reinterpret_cast<const float4 *>(vertsB); B.nverts = nB; B.cell_off = nullptr; B.cell_list = nullptr;
    pbp::SupMap ma, mb;
    auto make_map = [](const float *v4, int nv, pbp::SupMap &m) {
        std::vector<float> xyz(3 * nv);
        for (int i = 0; i < nv; i++) { xyz[3 * i] = v4[4 * i]; xyz[3 * i + 1] = v4[4 * i + 1]; xyz[3 * i + 2] = v4[4 * i + 2]; }
        m = pbp::build_support_map(xyz.data(), nv);
    };
    if (use_maps && shapeA == PBP_SHAPE_CONVEX) { make_map(vertsA, nA, ma); A.cell_off = ma.cell_off.data(); A.cell_list = ma.list.data(); }
    if (use_maps && shapeB == PBP_SHAPE_CONVEX) { make_map(vertsB, nB, mb); B.cell_off = mb.cell_off.data(); B.cell_list = mb.list.data(); }
    for (int64_t i = 0; i < n; i++) {
        const float3 v = make_float3(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]);
        pbp::GjkOut r = want_dist ? pbp::gjk_run<true>(A, B, T + 12 * i, v, margin, bound) : pbp::gjk_run<false>(A, B, T + 12 * i, v, margin, bound);
        hit[i] = r.hit;
        dist[i] = r.dist;
        iters[i] = r.iters;
    }
}

// witness variant: out [n][11] = dist, hit, far, pa(3), pb(3), n(3)... packed as 12 floats (last unused)
void emu_gjk_witness_batch(int shapeA, const float *parA, const float *vertsA, int nA, int shapeB, const float *parB,
                           const float *vertsB, int nB, const float *T, const float *v0, int64_t n, float ra, float rb,
                           float cutoff, float *out) {
    pbp::ShapeRef A, B;
    A.shape = shapeA; A.p0 = parA[0]; A.p1 = parA[1]; A.p2 = parA[2];
    A.verts = reinterpret_cast<const float4 *>(vertsA); A.nverts = nA; A.cell_off = nullptr; A.cell_list = nullptr;
    B.shape = shapeB; B.p0 = parB[0]; B.p1 = parB[1]; B.p2 = parB[2];
    B.verts = reinterpret_cast<const float4 *>(vertsB); B.nverts = nB; B.cell_off = nullptr; B.cell_list = nullptr;
    pbp::SupMap ma, mb;
    auto make_map = [](const float *v4, int nv, pbp::SupMap &m) {
        std::vector<float> xyz(3 * nv);
        for (int i = 0; i < nv; i++) { xyz[3 * i] = v4[4 * i]; xyz[3 * i + 1] = v4[4 * i + 1]; xyz[3 * i + 2] = v4[4 * i + 2]; }
        m = pbp::build_support_map(xyz.data(), nv);
    };
    if (shapeA == PBP_SHAPE_CONVEX) { make_map(vertsA, nA, ma); A.cell_off = ma.cell_off.data(); A.cell_list = ma.list.data(); }
    if (shapeB == PBP_SHAPE_CONVEX) { make_map(vertsB, nB, mb); B.cell_off = mb.cell_off.data(); B.cell_list = mb.list.data(); }
    for (int64_t i = 0; i < n; i++) {
        const float3 v = make_float3(v0[3 * i], v0[3 * i + 1], v0[3 * i + 2]);
        const pbp::WitnessOut w = pbp::gjk_run_witness(A, B, T + 12 * i, v, ra, rb, cutoff);
        float *o = out + 12 * i;
        o[0] = w.dist; o[1] = (float)w.hit; o[2] = (float)w.far;
        o[3] = w.pa.x; o[4] = w.pa.y; o[5] = w.pa.z; o[6] = w.pb.x; o[7] = w.pb.y; o[8] = w.pb.z;
        o[9] = w.n.x; o[10] = w.n.y; o[11] = w.n.z;
    }
}

// support-map statistics + exhaustive check against the full scan on random directions
// out: [0] list entries, [1] max candidates per cell, [2] mismatching directions (must be 0)
void emu_supmap_check(const float *verts4, int nv, const float *dirs, int64_t ndirs, int64_t *out) {
    std::vector<float> xyz(3 * nv);
    for (int i = 0; i < nv; i++) { xyz[3 * i] = verts4[4 * i]; xyz[3 * i + 1] = verts4[4 * i + 1]; xyz[3 * i + 2] = verts4[4 * i + 2]; }
    pbp::SupMap m = pbp::build_support_map(xyz.data(), nv);
    out[0] = (int64_t)m.list.size();
    out[1] = 0;
    for (int c = 0; c < pbp::SUPMAP_CELLS; c++) out[1] = std::max<int64_t>(out[1], m.cell_off[c + 1] - m.cell_off[c]);
    pbp::ShapeRef S;
    S.shape = PBP_SHAPE_CONVEX; S.verts = reinterpret_cast<const float4 *>(verts4); S.nverts = nv;
    S.cell_off = m.cell_off.data(); S.cell_list = m.list.data();
    int64_t bad = 0;
    for (int64_t i = 0; i < ndirs; i++) {
        const float3 d = make_float3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
        const float3 a = pbp::support_convex_map(S, d);
        const float3 b = pbp::support_convex(S.verts, nv, d);
        // equal support VALUE (ties may pick different vertices)
        if (pbp::dot3(a, d) != pbp::dot3(b, d)) bad++;
    }
    out[2] = bad;
}
}
