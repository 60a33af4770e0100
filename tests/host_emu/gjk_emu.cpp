// Host build of the engine's FP32 GJK (pyroboplan_b200/csrc/pbp_gjk.cuh) for CPU-side tests.
//
// TEST TOOLING ONLY: compiled by tests/ with g++ (no nvcc, no GPU), never loaded by the
// product.  It lets `pytest -m "not gpu"` exercise the exact float32 arithmetic of the device
// narrowphase (fmaf / IEEE sqrt / division are identical on host and device; only rsqrtf is
// emulated) against the float64 oracle on many random pairs.
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
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

// ------------------------------------------------------------------------------------------
// Surface semantics (pbp_surface.cuh) on the host: one "lane" walks every warp-level loop.
// ------------------------------------------------------------------------------------------
static std::string g_emu_error;
static int emu_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_emu_error = buf;
    return code;
}
#define PBP_HOST_FAIL emu_fail
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline void __syncwarp() {}
#include "../../pyroboplan_b200/csrc/pbp_host_tables.h"
#include "../../pyroboplan_b200/csrc/pbp_surface.cuh"

extern "C" {

const char *emu_last_error() { return g_emu_error.c_str(); }

// the triangle stage's closed-form distance on n triangle pairs (A, B: [n][9] float32)
void emu_tri_tri(const float *A, const float *B, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) {
        float3 a[3], b[3];
        for (int k = 0; k < 3; k++) {
            a[k] = make_float3(A[9 * i + 3 * k], A[9 * i + 3 * k + 1], A[9 * i + 3 * k + 2]);
            b[k] = make_float3(B[9 * i + 3 * k], B[9 * i + 3 * k + 1], B[9 * i + 3 * k + 2]);
        }
        out[i] = pbp::tri_tri_distance(a, b);
    }
}

// The item setup's hit certificate for n placements of the caller's pair: out[i] bit 0 = the inner balls are within
// the padding, bit 1 = the centre-line test rules an enclosure out (always set for pairs without an enclosure flag).
// The kernel publishes a hit when both are set.
int emu_ball_certificate(const pbp_model_desc *d, int pair, const float *T, int64_t n, float padding, int32_t *out) {
    static HostTables H;
    static const void *h_key = nullptr;
    if (h_key != (const void *)d->verts) {
        H = HostTables();
        const int rc = build_host_tables(d, H);
        if (rc) return rc;
        h_key = (const void *)d->verts;
    }
    int p = -1;
    for (int k = 0; k < d->npairs; k++)
        if (H.pair_orig[k] == pair) p = k;
    if (p < 0) return emu_fail(-1, "no such pair");
    SurfTables tab{};
    tab.hverts = H.verts.data(); tab.hcells = H.sup_cells.data(); tab.hlists = H.sup_lists.data();
    const BpPair pr = H.bppairs[p];
    const DevGeom &GA = H.geoms[pr.a];
    const DevGeom &GB = H.geoms[pr.b];
    const int fl = H.has_tris ? (H.pair_flags[p] & PAIR_ENCLOSE) : 0;
    for (int64_t i = 0; i < n; i++) {
        const float *Ti = T + 12 * i;
        const bool balls = inner_balls_hit(GA, GB, Ti, padding);
        const bool pokes = !fl || enclosure_quick_pokes(fl, GA, GB, hull_shape(tab, GA), hull_shape(tab, GB), Ti, padding);
        out[i] = (balls ? 1 : 0) | (pokes ? 2 : 0);
    }
    return 0;
}

// What the engine answers for n placements T [n][12] (b in a's frame) of the caller's pair `pair`:
// the narrowphase's GJK on the cores (boolean) / hulls (distance), then the refine kernel's lane and
// warp decisions.  mode 0: out_hit[i] = verdict at `padding`; mode 2: out_dist[i] = surface distance
// of the pair (the state's running minimum starts at +inf).  out_path[i]: 0 decided by the GJK alone,
// 1 lane level (hull / core runs, quick enclosure tests), 2 plane walk, 3 triangle stage.
int emu_surface_items(const pbp_model_desc *d, int pair, const float *T, int64_t n, float padding, int mode, int32_t *out_hit, float *out_dist,
                      int32_t *out_path) {
    // the Python driver keeps the flattened arrays of a model alive: rebuild the tables only for a new one
    static HostTables H;
    static const void *h_key = nullptr;
    if (h_key != (const void *)d->verts) {
        H = HostTables();
        const int rc = build_host_tables(d, H);
        if (rc) return rc;
        h_key = (const void *)d->verts;
    }
    int p = -1;
    for (int k = 0; k < d->npairs; k++)
        if (H.pair_orig[k] == pair) p = k;
    if (p < 0) return emu_fail(-1, "no such pair");
    DevModel M{};
    M.npairs = d->npairs; M.ngeoms = d->ngeoms;
    M.geoms = H.geoms.data(); M.bppairs = H.bppairs.data(); M.pair_enclose = H.pair_flags.data();
    M.has_tris = H.has_tris ? 1 : 0;
    std::vector<float4> planes((size_t)std::max(d->nplanes, 1));
    for (int i = 0; i < d->nplanes; i++) planes[i] = make_float4(d->planes[4 * i], d->planes[4 * i + 1], d->planes[4 * i + 2], d->planes[4 * i + 3]);
    SurfTables tab{};
    tab.hverts = H.verts.data(); tab.hcells = H.sup_cells.data(); tab.hlists = H.sup_lists.data();
    tab.kverts = H.has_tris ? H.kverts.data() : H.verts.data();
    tab.kcells = H.has_tris ? H.ksup_cells.data() : H.sup_cells.data();
    tab.klists = H.has_tris ? H.ksup_lists.data() : H.sup_lists.data();
    tab.planes = planes.data(); tab.mverts = H.mverts.data(); tab.tris = H.tris.data(); tab.tverts = H.tverts.data();
    const BpPair pr = H.bppairs[p];
    const DevGeom &GA = H.geoms[pr.a];
    const DevGeom &GB = H.geoms[pr.b];
    const int fl = H.pair_flags[p];
    const float radii = GA.core_radius + GB.core_radius;
    const float margin = padding + radii;
    std::vector<uint16_t> la(TRI_LIST_CAP), lb(TRI_LIST_CAP);
    for (int64_t i = 0; i < n; i++) {
        const float *Ti = T + 12 * i;
        int hit = 0, path = 0;
        float best = INFINITY;
        auto on_hit = [&](int, int32_t, int64_t) { hit = 1; };
        auto on_value = [&](int, int32_t, int64_t, float dd) { best = std::min(best, dd); if (dd <= padding) hit = 1; };
        auto decided = [&](int32_t) { return false; };
        LaneVerdict v;
        v.status = ST_NONE; v.dist = 0.f; v.U = 0.f; v.n = make_float3(0.f, 0.f, 0.f); v.have_n = 0;
        if (mode == MODE_DIST) {
            const ShapeRef A = hull_shape(tab, GA), B = hull_shape(tab, GB);
            const GjkOut r = gjk_run<true>(A, B, Ti, gjk_start_direction(GA, GB, Ti), margin, INFINITY);
            if (fl & PAIR_SURFACE) {
                path = 1;
                v = decide_distance(tab, fl, GA, GB, Ti, INFINITY);
                if (v.status == ST_VALUE) on_value(p, 0, 0, fmaxf(v.dist - radii, 0.f));
                if (v.status == ST_WALK) path = 2;
                if (v.status == ST_TRIANGLES) path = 3;
                settle_warp_items<MODE_DIST>(M, tab, v, p, 0, (int64_t)0, Ti, padding, la.data(), lb.data(), 0, on_hit, on_value, decided);
            } else {
                on_value(p, 0, 0, fmaxf(r.dist - radii, 0.f));
            }
        } else {
            // as k_narrowphase runs it: boolean GJK on the cores, leaving as undecided as soon as the cores are
            // certainly apart beyond the margin without being beyond the cut; the direction it ends on travels on
            const ShapeRef A = core_shape(tab, GA), B = core_shape(tab, GB);
            GjkState gs;
            GjkOut r;
            gjk_init(gs, gjk_start_direction(GA, GB, Ti));
            const float cut = margin + pair_eps(GA, GB);
            for (;;) {
                if (gjk_support_phase<false, false, true>(A, B, Ti, gs, margin, cut, r)) break;
                if (gjk_simplex_phase<false>(gs, margin, r)) break;
            }
            bool parked = false, doubt = false;
            if (fl & PAIR_SURFACE) {
                if (r.hit && (fl & PAIR_NOCORE)) doubt = true;
                else if (r.hit && (fl & PAIR_ENCLOSE)) parked = true;
                else if (!r.hit && !r.far) doubt = true;
            }
            if (!parked && !doubt) {
                hit = r.hit;
            } else {
                path = 1;
                const bool warm = doubt && !r.hit && !(fl & PAIR_NOCORE);
                v = decide_boolean(tab, fl, parked, GA, GB, Ti, padding, warm ? gs.v : make_float3(0.f, 0.f, 0.f));
                if (v.status == ST_HIT) hit = 1;
                if (v.status == ST_WALK) path = 2;
                if (v.status == ST_TRIANGLES) path = 3;
                settle_warp_items<MODE_BOOL>(M, tab, v, p, 0, (int64_t)0, Ti, padding, la.data(), lb.data(), 0, on_hit, on_value, decided);
            }
        }
        out_hit[i] = hit;
        out_dist[i] = best;
        out_path[i] = path;
    }
    return 0;
}
}
