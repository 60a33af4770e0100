// Host-side compilation of a model description (pbp_model_desc) into the kernels' tables.  Plain C++,
// no CUDA calls: included by pbp_engine.cu (the product) and by the host build of the device headers
// the CPU tests use (tests/host_emu).  The including file defines PBP_HOST_FAIL(code, fmt, ...) -- it
// records the message and returns `code`.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "pbp_model.cuh"
#include "pbp_supmap.h"

#ifndef PBP_HOST_FAIL
#error "define PBP_HOST_FAIL(code, fmt, ...) before including pbp_host_tables.h"
#endif

namespace pbp {
#ifndef PBP_PAIR_FLAGS_DEFINED
#define PBP_PAIR_FLAGS_DEFINED
// per-pair flags (DevModel::pair_enclose, internal pair order); used by the kernels in pbp_surface.cuh
enum : uint8_t {
    PAIR_ENCLOSE_B_IN_A = 1,   // `second` could lie wholly inside `first`, and `first` is a surface
    PAIR_ENCLOSE_A_IN_B = 2,   // `first` could lie wholly inside `second`, and `second` is a surface
    PAIR_ENCLOSE = 3,
    PAIR_SURFACE = 4,          // a geometry of the pair carries triangles
    PAIR_NOCORE = 8,           // ... and has no inner core: a hull hit proves nothing, the triangles decide every one
};
#endif
constexpr int BP_MAX_SAVES_HOST = 16;
}  // namespace pbp

using namespace pbp;

struct HostTables {
    std::vector<DevJoint> joints;
    std::vector<DevGeom> geoms;
    std::vector<BpGeom> bpgeoms;
    std::vector<BpBox> bpboxes;
    std::vector<float4> verts;
    std::vector<uint16_t> sup_cells, sup_lists;
    std::vector<float4> kverts;                    // inner cores (empty: no geometry is a surface, the hull tables serve)
    std::vector<uint16_t> ksup_cells, ksup_lists;
    std::vector<float4> mverts;                    // mesh vertices and triangles of surface geometries
    std::vector<int4> tris;
    std::vector<float4> tverts;   // the three vertices of every triangle, packed in triangle order (no index indirection on the device)
    std::vector<uint8_t> pair_flags;               // PAIR_* per internal pair
    bool has_tris = false;
    std::vector<int32_t> jgeoms;
    std::vector<BpPair> bppairs;
    std::vector<BpGroup> groups;
    std::vector<int32_t> pair_orig;
    std::vector<PairPath> paths;
    std::vector<uint8_t> pathops;
    std::vector<int32_t> order;
    int nmoving = 0, nsave = 0;
};

inline int build_host_tables(const pbp_model_desc *d, HostTables &H) {
    // ---- joints (+ which transforms the broadphase must keep for non-adjacent children)
    H.joints.assign(d->njoints, DevJoint{});
    auto &joints = H.joints;
    for (int j = 0; j < d->njoints; j++) {
        DevJoint &J = joints[j];
        memset(&J, 0, sizeof J);
        memcpy(J.P, d->joint_placement + 12 * j, sizeof J.P);
        memcpy(J.axis, d->joint_axis + 3 * j, sizeof J.axis);
        J.parent = d->joint_parent[j];
        J.type = d->joint_type[j];
        J.idx_q = d->joint_idx_q[j];
        J.axis_code = 0;
        for (int k = 0; k < 3; k++) {
            const float a = J.axis[k], b = J.axis[(k + 1) % 3], c = J.axis[(k + 2) % 3];
            if (fabsf(fabsf(a) - 1.0f) < 1e-7f && b == 0.0f && c == 0.0f) J.axis_code = k + 1;
        }
        J.parent_slot = -1;
        J.save_slot = -1;
        if (j > 0 && (J.parent < 0 || J.parent >= j)) return (PBP_HOST_FAIL(PBP_ERR_ARG, "joint %d: parent %d not topologically ordered", j, J.parent));
        if (j > 0 && (J.idx_q < 0 || J.idx_q >= d->nq)) return (PBP_HOST_FAIL(PBP_ERR_ARG, "joint %d: idx_q %d out of range", j, J.idx_q));
        if (j > 0 && J.type != PBP_JOINT_REVOLUTE && J.type != PBP_JOINT_PRISMATIC) return (PBP_HOST_FAIL(PBP_ERR_ARG, "joint %d: unsupported type %d", j, J.type));
    }
    int nsave = 0;
    for (int j = 1; j < d->njoints; j++) {
        DevJoint &J = joints[j];
        if (J.parent == 0) J.parent_slot = -2;
        else if (J.parent == j - 1) J.parent_slot = -1;
        else {
            if (joints[J.parent].save_slot < 0) joints[J.parent].save_slot = nsave++;
            J.parent_slot = joints[J.parent].save_slot;
        }
    }
    if (nsave > BP_MAX_SAVES_HOST) return (PBP_HOST_FAIL(PBP_ERR_CAPACITY, "kinematic tree has %d branch points (max %d)", nsave, BP_MAX_SAVES_HOST));

    // ---- geometries: narrowphase view, broadphase proxies, static boxes, padded vertex table
    H.geoms.assign(std::max(d->ngeoms, 1), DevGeom{});
    auto &geoms = H.geoms;
    H.bpgeoms.assign(std::max(d->ngeoms, 1), BpGeom{});
    auto &bpgeoms = H.bpgeoms;
    auto &bpboxes = H.bpboxes;
    auto &verts = H.verts;
    auto &sup_cells = H.sup_cells;
    auto &sup_lists = H.sup_lists;
    auto &jgeoms = H.jgeoms;
    int nmoving = 0;
    for (int g = 0; g < d->ngeoms; g++) {
        DevGeom &G = geoms[g];
        BpGeom &B = bpgeoms[g];
        memset(&G, 0, sizeof G);
        memset(&B, 0, sizeof B);
        G.map_off = -1;
        G.list_off = 0;
        memcpy(G.P, d->geom_placement + 12 * g, sizeof G.P);
        memcpy(G.par, d->geom_params + 4 * g, sizeof G.par);
        const float *outer = d->geom_outer + 4 * g, *inner = d->geom_inner + 4 * g;
        for (int i = 0; i < 3; i++)
            if (fabsf(outer[i] - inner[i]) > 1e-6f) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: geom_inner must share the centre of geom_outer", g));
        if (!(inner[3] <= outer[3])) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: inscribed radius exceeds enclosing radius", g));
        memcpy(G.centre, outer, sizeof G.centre);
        G.parent = d->geom_parent[g];
        G.shape = d->geom_shape[g];
        if (G.parent < 0 || G.parent >= d->njoints) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad parent joint", g));
        if (G.shape < PBP_SHAPE_SPHERE || G.shape > PBP_SHAPE_CONVEX) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad shape %d", g, G.shape));
        G.core_radius = (G.shape == PBP_SHAPE_SPHERE || G.shape == PBP_SHAPE_CAPSULE) ? G.par[0] : 0.f;
        // proxy centre in the parent-joint frame (= world frame for static geometries)
        for (int i = 0; i < 3; i++) B.c[i] = G.P[3 * i] * outer[0] + G.P[3 * i + 1] * outer[1] + G.P[3 * i + 2] * outer[2] + G.P[9 + i];
        B.r_out = outer[3];
        B.r_in = inner[3];
        G.r_out = outer[3];
        if (d->geom_inner_balls) {
            for (int k = 0; k < N_INNER_BALLS; k++) {
                const float *b = d->geom_inner_balls + (size_t)(g * N_INNER_BALLS + k) * 4;
                if (!(b[3] > 0.f)) continue;
                memcpy(G.balls[G.n_balls++], b, 4 * sizeof(float));
            }
        }
        if (d->geom_capsule) {   // the capsule's segment doubles as the shape's "long axis" for the GJK start direction
            const float *cp = d->geom_capsule + 8 * g;
            const float du[3] = {cp[3] - cp[0], cp[4] - cp[1], cp[5] - cp[2]};
            if (cp[6] > 0.f && du[0] * du[0] + du[1] * du[1] + du[2] * du[2] > 1e-10f) {
                G.has_axis = 1;
                for (int i = 0; i < 3; i++) { G.axis0[i] = cp[i]; G.axis1[i] = cp[3 + i]; }
            }
        }
        if (d->geom_capsule && d->geom_capsule[8 * g + 7] != 0.f) {
            const float *cp = d->geom_capsule + 8 * g;
            B.has_cap = 1;
            B.cap_r = cp[6];
            for (int i = 0; i < 3; i++) {
                B.cap0[i] = G.P[3 * i] * cp[0] + G.P[3 * i + 1] * cp[1] + G.P[3 * i + 2] * cp[2] + G.P[9 + i];
                B.cap1[i] = G.P[3 * i] * cp[3] + G.P[3 * i + 1] * cp[4] + G.P[3 * i + 2] * cp[5] + G.P[9 + i];
            }
        }
        B.slot = G.parent == 0 ? -1 : nmoving++;
        B.box = -1;
        if (G.parent == 0 && G.shape == PBP_SHAPE_BOX) {
            BpBox bx;
            memset(&bx, 0, sizeof bx);
            memcpy(bx.R, G.P, sizeof bx.R);
            memcpy(bx.t, G.P + 9, sizeof bx.t);
            memcpy(bx.h, G.par, sizeof bx.h);
            B.box = (int)bpboxes.size();
            bpboxes.push_back(bx);
        }
        if (G.shape == PBP_SHAPE_CONVEX) {
            const int off = d->geom_vert_off[g], cnt = d->geom_vert_cnt[g];
            if (cnt < 1 || off < 0 || off + cnt > d->nverts) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad vertex range", g));
            G.vert_off = (int)verts.size();
            for (int i = 0; i < cnt; i++) verts.push_back(make_float4(d->verts[3 * (off + i)], d->verts[3 * (off + i) + 1], d->verts[3 * (off + i) + 2], 0.f));
            while (verts.size() % 4) verts.push_back(verts.back());
            G.vert_cnt = (int)verts.size() - G.vert_off;
            // exact support map of this hull (shared by geometries with identical vertex data)
            int same = -1;
            for (int h = 0; h < g && same < 0; h++)
                if (geoms[h].shape == PBP_SHAPE_CONVEX && d->geom_vert_cnt[h] == cnt &&
                    memcmp(d->verts + 3 * d->geom_vert_off[h], d->verts + 3 * off, (size_t)cnt * 12) == 0)
                    same = h;
            if (same >= 0) {
                G.map_off = geoms[same].map_off;
                G.list_off = geoms[same].list_off;
            } else if (cnt <= 65535) {
                const SupMap sm = build_support_map(d->verts + 3 * off, cnt);
                if (sm.list.size() <= 65535) {
                    G.map_off = (int)sup_cells.size();
                    G.list_off = (int)sup_lists.size();
                    sup_cells.insert(sup_cells.end(), sm.cell_off.begin(), sm.cell_off.end());
                    sup_lists.insert(sup_lists.end(), sm.list.begin(), sm.list.end());
                }
            }
        }
    }
    while (sup_cells.empty() || sup_cells.size() % 8) sup_cells.push_back(0);
    while (sup_lists.empty() || sup_lists.size() % 8) sup_lists.push_back(0);
    // ---- surface geometries: inner cores (second set of GJK tables), mesh vertices, triangles
    const bool have_tris = d->tris && d->geom_tri_off && d->geom_tri_cnt && d->mesh_verts && d->geom_mvert_off && d->ntris > 0;
    const bool have_cores = d->core_verts && d->geom_core_off && d->geom_core_cnt && d->ncore > 0;
    for (int g = 0; g < d->ngeoms; g++) {
        DevGeom &G = geoms[g];
        G.kvert_off = G.vert_off; G.kvert_cnt = G.vert_cnt; G.kmap_off = G.map_off; G.klist_off = G.list_off;
        G.tri_off = 0; G.tri_cnt = 0; G.mvert_off = 0; G.core_eps = 0.f;
        G.surface_tol = d->geom_surface_tol ? d->geom_surface_tol[g] : 0.f;
        G.plane_off = d->geom_plane_off ? d->geom_plane_off[g] : 0;
        G.plane_cnt = d->geom_plane_cnt ? d->geom_plane_cnt[g] : 0;
        G.plane_hull_cnt = d->geom_plane_hull_cnt ? d->geom_plane_hull_cnt[g] : G.plane_cnt;
        if (G.plane_hull_cnt < 0 || G.plane_hull_cnt > G.plane_cnt) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad hull plane count", g));
    }
    if (have_tris) {
        for (int g = 0; g < d->ngeoms; g++) {
            const int cnt = d->geom_tri_cnt[g];
            if (cnt == 0) continue;
            DevGeom &G = geoms[g];
            if (G.shape != PBP_SHAPE_CONVEX) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: triangles on a primitive", g));
            const int off = d->geom_tri_off[g], mo = d->geom_mvert_off[g];
            if (cnt < 0 || cnt > 65535 || off < 0 || off + cnt > d->ntris || mo < 0 || mo > d->nmverts)
                return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad triangle range", g));
            int maxv = -1;
            G.tri_off = (int)H.tris.size();
            G.tri_cnt = cnt;
            for (int t = 0; t < cnt; t++) {
                const int32_t *ix = d->tris + 3 * (size_t)(off + t);
                for (int k = 0; k < 3; k++) {
                    if (ix[k] < 0 || mo + ix[k] >= d->nmverts) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: triangle %d: vertex index out of range", g, t));
                    maxv = std::max(maxv, ix[k]);
                }
                H.tris.push_back(make_int4(ix[0], ix[1], ix[2], 0));
                for (int k = 0; k < 3; k++)
                    H.tverts.push_back(make_float4(d->mesh_verts[3 * (size_t)(mo + ix[k])], d->mesh_verts[3 * (size_t)(mo + ix[k]) + 1],
                                                   d->mesh_verts[3 * (size_t)(mo + ix[k]) + 2], 0.f));
            }
            G.mvert_off = (int)H.mverts.size();
            for (int i = 0; i <= maxv; i++)
                H.mverts.push_back(make_float4(d->mesh_verts[3 * (size_t)(mo + i)], d->mesh_verts[3 * (size_t)(mo + i) + 1], d->mesh_verts[3 * (size_t)(mo + i) + 2], 0.f));
            H.has_tris = true;
        }
    }
    if (H.has_tris) {
        // second set of GJK tables: the core of every surface geometry that has one, a copy of the hull otherwise
        struct Built { const float *src; int cnt; int g; };
        std::vector<Built> built;
        for (int g = 0; g < d->ngeoms; g++) {
            DevGeom &G = geoms[g];
            if (G.shape != PBP_SHAPE_CONVEX) continue;
            const bool core = have_cores && G.tri_cnt > 0 && d->geom_core_cnt[g] > 0;
            const float *src = core ? d->core_verts + 3 * (size_t)d->geom_core_off[g] : d->verts + 3 * (size_t)d->geom_vert_off[g];
            const int cnt = core ? d->geom_core_cnt[g] : d->geom_vert_cnt[g];
            if (core && (d->geom_core_off[g] < 0 || d->geom_core_off[g] + cnt > d->ncore)) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad core vertex range", g));
            if (core) G.core_eps = d->geom_core_eps ? d->geom_core_eps[g] : 0.f;
            if (!(G.core_eps >= 0.f) || G.core_eps > 1e3f) return (PBP_HOST_FAIL(PBP_ERR_ARG, "geometry %d: bad core_eps", g));
            int same = -1;
            for (const Built &b : built)
                if (b.cnt == cnt && memcmp(b.src, src, (size_t)cnt * 12) == 0) { same = b.g; break; }
            if (same >= 0) {
                G.kvert_off = geoms[same].kvert_off; G.kvert_cnt = geoms[same].kvert_cnt;
                G.kmap_off = geoms[same].kmap_off; G.klist_off = geoms[same].klist_off;
                continue;
            }
            G.kvert_off = (int)H.kverts.size();
            for (int i = 0; i < cnt; i++) H.kverts.push_back(make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f));
            while (H.kverts.size() % 4) H.kverts.push_back(H.kverts.back());
            G.kvert_cnt = (int)H.kverts.size() - G.kvert_off;
            G.kmap_off = -1;
            G.klist_off = 0;
            if (!core && G.map_off >= 0) {
                // the hull's own map, copied (offsets are relative to the hull's first vertex)
                const int h_cells = SUPMAP_CELLS + 1;
                G.kmap_off = (int)H.ksup_cells.size();
                G.klist_off = (int)H.ksup_lists.size();
                const int first = sup_cells[G.map_off], last = sup_cells[G.map_off + SUPMAP_CELLS];
                for (int c = 0; c < h_cells; c++) H.ksup_cells.push_back((uint16_t)(sup_cells[G.map_off + c]));
                for (int i = first; i < last + 4; i++) H.ksup_lists.push_back(sup_lists[G.list_off + i]);
            } else if (cnt <= 65535) {
                const SupMap sm = build_support_map(src, cnt);
                if (sm.list.size() <= 65535) {
                    G.kmap_off = (int)H.ksup_cells.size();
                    G.klist_off = (int)H.ksup_lists.size();
                    H.ksup_cells.insert(H.ksup_cells.end(), sm.cell_off.begin(), sm.cell_off.end());
                    H.ksup_lists.insert(H.ksup_lists.end(), sm.list.begin(), sm.list.end());
                }
            }
            built.push_back(Built{src, cnt, g});
        }
        while (H.kverts.empty() || H.kverts.size() % 4) H.kverts.push_back(make_float4(0, 0, 0, 0));
        while (H.ksup_cells.empty() || H.ksup_cells.size() % 8) H.ksup_cells.push_back(0);
        while (H.ksup_lists.empty() || H.ksup_lists.size() % 8) H.ksup_lists.push_back(0);
        if (H.mverts.empty()) H.mverts.push_back(make_float4(0, 0, 0, 0));
    }
    // moving geometries grouped by parent joint (the broadphase places them right after that joint)
    for (int j = 1; j < d->njoints; j++) {
        joints[j].geom_begin = (int)jgeoms.size();
        for (int g = 0; g < d->ngeoms; g++)
            if (geoms[g].parent == j) jgeoms.push_back(g);
        joints[j].geom_end = (int)jgeoms.size();
    }
    if (verts.empty()) verts.push_back(make_float4(0, 0, 0, 0));
    if (bpboxes.empty()) bpboxes.push_back(BpBox{});
    if (jgeoms.empty()) jgeoms.push_back(0);
    H.nmoving = nmoving;
    H.nsave = nsave;

    // ---- pairs: internal order (moving-moving first, then one contiguous group per static
    //      geometry), broadphase recipe, joint path for the item setup, bin order
    const int np1 = std::max(d->npairs, 1);
    struct PairKey { int cls, sidx, orig; };
    std::vector<PairKey> keys(d->npairs);
    for (int p = 0; p < d->npairs; p++) {
        const int a = d->pairs[2 * p], b = d->pairs[2 * p + 1];
        if (a < 0 || b < 0 || a >= d->ngeoms || b >= d->ngeoms || a == b) return (PBP_HOST_FAIL(PBP_ERR_ARG, "pair %d: bad geometry ids", p));
        const bool sa = bpgeoms[a].slot < 0, sb = bpgeoms[b].slot < 0;
        PairKey k{GRP_GENERIC, 0, p};
        if (!sa && !sb) k = PairKey{GRP_MOVING, 0, p};
        else if (sa != sb) {
            const int st = sa ? a : b;
            if (bpgeoms[st].box >= 0) k = PairKey{GRP_STATIC_BOX, bpgeoms[st].box, p};
            else k = PairKey{GRP_STATIC_POINT, st, p};
        }
        keys[p] = k;
    }
    std::stable_sort(keys.begin(), keys.end(), [](const PairKey &x, const PairKey &y) {
        return x.cls != y.cls ? x.cls < y.cls : x.sidx < y.sidx;
    });
    H.bppairs.assign(np1, BpPair{});
    auto &bppairs = H.bppairs;
    auto &groups = H.groups;
    H.pair_orig.assign(np1, 0);
    auto &pair_orig = H.pair_orig;
    H.paths.assign(np1, PairPath{});
    auto &paths = H.paths;
    auto &pathops = H.pathops;
    H.order.assign(np1, 0);
    H.pair_flags.assign(np1, 0);
    auto &order = H.order;
    std::vector<long> cost(np1, 0);
    auto chain_to_root = [&](int j) {  // joints from the root down to j (empty for the universe)
        std::vector<int> c;
        for (; j > 0; j = joints[j].parent) c.push_back(j);
        std::reverse(c.begin(), c.end());
        return c;
    };
    for (int p = 0; p < d->npairs; p++) {
        const PairKey &key = keys[p];
        const int a = d->pairs[2 * key.orig], b = d->pairs[2 * key.orig + 1];
        pair_orig[p] = key.orig;
        if (groups.empty() || groups.back().cls != key.cls || (key.cls != GRP_MOVING && key.cls != GRP_GENERIC && groups.back().sidx != key.sidx))
            groups.push_back(BpGroup{key.cls, key.sidx, p, p});
        groups.back().end = p + 1;
        BpPair &P = bppairs[p];
        memset(&P, 0, sizeof P);
        P.a = (uint8_t)a;
        P.b = (uint8_t)b;
        P.slot_a = (int16_t)bpgeoms[a].slot;
        P.slot_b = (int16_t)bpgeoms[b].slot;
        const int sha = geoms[a].shape, shb = geoms[b].shape;
        if (sha == PBP_SHAPE_SPHERE && shb == PBP_SHAPE_SPHERE) {
            P.kind = BP_EXACT_SPHERES;
            P.r_out = P.r_in = geoms[a].par[0] + geoms[b].par[0];
        } else if (bpgeoms[a].box >= 0) {
            P.kind = BP_BOX_A;
            P.r_out = bpgeoms[b].r_out;
            P.r_in = bpgeoms[b].r_in;
        } else if (bpgeoms[b].box >= 0) {
            P.kind = BP_BOX_B;
            P.r_out = bpgeoms[a].r_out;
            P.r_in = bpgeoms[a].r_in;
        } else {
            P.kind = BP_SPHERES;
            P.r_out = bpgeoms[a].r_out + bpgeoms[b].r_out;
            P.r_in = bpgeoms[a].r_in + bpgeoms[b].r_in;
        }
        // surface semantics: a pair where one shape could be enclosed by the other's surface is never
        // "certainly colliding" by its inscribed spheres (the enclosed placement is the free one)
        uint8_t pflags = 0;
        if (d->pair_enclose && d->planes && d->nplanes > 0) pflags |= d->pair_enclose[key.orig] & PAIR_ENCLOSE;
        if (geoms[a].tri_cnt > 0 || geoms[b].tri_cnt > 0) pflags |= PAIR_SURFACE;
        if ((geoms[a].tri_cnt > 0 && !(have_cores && d->geom_core_cnt[a] > 0)) || (geoms[b].tri_cnt > 0 && !(have_cores && d->geom_core_cnt[b] > 0)))
            pflags |= PAIR_NOCORE;
        if (((pflags & PAIR_ENCLOSE_B_IN_A) && geoms[a].plane_cnt == 0) || ((pflags & PAIR_ENCLOSE_A_IN_B) && geoms[b].plane_cnt == 0))
            return (PBP_HOST_FAIL(PBP_ERR_ARG, "pair %d: enclosure flag on a geometry without planes", key.orig));
        H.pair_flags[p] = pflags;
        // a pair where one shape could be enclosed by the other's surface, or whose surface has no inner core,
        // is never "certainly colliding" by its inscribed spheres
        if ((pflags & (PAIR_ENCLOSE | PAIR_NOCORE)) && P.kind != BP_EXACT_SPHERES) P.r_in = -1e30f;
        const std::vector<int> ca = chain_to_root(geoms[a].parent), cb = chain_to_root(geoms[b].parent);
        size_t common = 0;
        while (common < ca.size() && common < cb.size() && ca[common] == cb[common]) common++;
        paths[p].off = (int32_t)pathops.size();
        paths[p].n_trunk = (uint8_t)common;
        paths[p].n_a = (uint8_t)(ca.size() - common);
        paths[p].n_b = (uint8_t)(cb.size() - common);
        for (size_t i = 0; i < ca.size(); i++) pathops.push_back((uint8_t)ca[i]);
        for (size_t i = common; i < cb.size(); i++) pathops.push_back((uint8_t)cb[i]);
        cost[p] = 8 + geoms[a].vert_cnt + geoms[b].vert_cnt;
        order[p] = p;
    }
    if (groups.empty()) groups.push_back(BpGroup{GRP_GENERIC, 0, 0, 0});
    if (pathops.empty()) pathops.push_back(0);
    while (pathops.size() % 4) pathops.push_back(0);
    std::stable_sort(order.begin(), order.begin() + d->npairs, [&](int x, int y) { return cost[x] > cost[y]; });

    return 0;
}

