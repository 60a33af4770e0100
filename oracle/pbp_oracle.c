/*
 * pbp_oracle.c -- float64 CPU restatement of the reference's collision hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (pyroboplan_b200/) may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, and only as the checker or the reported CPU baseline.
 *
 * What is restated, and from where
 * --------------------------------
 * The reference's arithmetic for this path lives in un-vendored third-party code:
 * Pinocchio (pin == 3.7.0, /root/reference/pyproject.toml:10) and coal (transitive, unpinned).
 * Neither is installable here, so this file restates their published algorithms and anchors
 * on the reference's own call sites:
 *   - pinocchio.computeCollisions(model, data, cmodel, cdata, q, stop_at_first=True)
 *       /root/reference/src/pyroboplan/core/utils.py:48-51 and :125-130
 *     = forward kinematics  oMi[j] = oMi[parent] * jointPlacement[j] * T_j(q_j),
 *       geometry placement  oMg[g] = oMi[parentJoint(g)] * placement(g),
 *       then, pair by pair in list order, coal collide(): collision <=> distance <=
 *       security_margin, returning at the first colliding pair.
 *   - pinocchio.computeDistances  (core/utils.py:86): min over pairs of the pair distance.
 *   - discretize_joint_space_path  /root/reference/src/pyroboplan/planning/utils.py:114-140
 *   - has_collision_free_path      /root/reference/src/pyroboplan/planning/utils.py:51-111
 *   - check_collisions_along_path  /root/reference/src/pyroboplan/core/utils.py:90-132
 * Pair distance: GJK (Gilbert-Johnson-Keerthi) on support functions, float64, run to
 * convergence (exact for polytopes), with spheres/capsules handled as core + radius.  URDF
 * meshes are SURFACES, as Pinocchio's coal BVHModels are: the convex hull of a mesh is only the
 * first stage -- a pair whose hulls are within the padding is decided by exact triangle tests
 * (oracle_soup_pair_distance, end of file; DESIGN.md "Mesh semantics"), so a shape wholly inside a
 * mesh does not collide with it.  Without triangle tables (oracle_model.tris == NULL) meshes are
 * convex solids.  The verdict path reports penetrating pairs with distance 0;
 * oracle_pair_signed_distance restates coal's penetration depth (EPA) on the hulls for the
 * distance / avoidance rows.
 *
 * PARITY UNPINNED beyond the reference's own known-answer vectors: no output of Pinocchio / coal
 * is available here.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against every known-answer vector
 * the reference's tests hold for the path and verifies, independently of GJK,
 * the separating-plane / common-point certificate this file emits for each pair result.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

enum { SH_SPHERE = 0, SH_BOX = 1, SH_CYLINDER = 2, SH_CAPSULE = 3, SH_CONVEX = 4 };
enum { JT_FIXED = 0, JT_REVOLUTE = 1, JT_PRISMATIC = 2 };

typedef struct {
    int32_t nq, njoints, ngeoms, npairs, nverts;
    const int32_t *joint_parent;   /* [njoints] */
    const int32_t *joint_type;     /* [njoints] */
    const int32_t *joint_idx_q;    /* [njoints] */
    const double *joint_placement; /* [njoints][12]  R row-major (9) then t (3) */
    const double *joint_axis;      /* [njoints][3] unit */
    const int32_t *geom_parent;    /* [ngeoms] parent joint */
    const int32_t *geom_shape;     /* [ngeoms] */
    const double *geom_placement;  /* [ngeoms][12] */
    const double *geom_params;     /* [ngeoms][4] sphere r | box hx hy hz | cyl r hl | capsule r hl */
    const int32_t *geom_vert_off;  /* [ngeoms] */
    const int32_t *geom_vert_cnt;  /* [ngeoms] */
    const double *geom_bsphere;    /* [ngeoms][4] bounding sphere cx cy cz r in geometry frame */
    const double *geom_isphere;    /* [ngeoms][4] inscribed sphere cx cy cz r in geometry frame */
    const double *verts;           /* [nverts][3] */
    const int32_t *pairs;          /* [npairs][2] */
    /* optional (NULL: every geometry is a solid): surface triangles of the mesh geometries, the
     * reference's semantics for URDF meshes (see "triangle-soup semantics" at the end of the file) */
    const double *tris;            /* [ntris][9] vertex coordinates, geometry frame */
    const int32_t *tri_off;        /* [ngeoms] */
    const int32_t *tri_cnt;        /* [ngeoms] 0 = solid primitive / solid hull */
} oracle_model;

typedef struct {
    int64_t pair_tests;    /* pairs that reached the narrowphase */
    int64_t gjk_iters;     /* total GJK iterations */
    int64_t support_calls; /* total single-shape support evaluations */
    int64_t support_verts; /* total hull vertices scanned by support evaluations */
    int64_t culled;        /* pairs rejected by the bounding-sphere test */
} oracle_counters;

/* ------------------------------------------------------------------ small vector helpers */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static inline void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
/* T = [R(9) | t(3)] */
static void se3_mul(const double *A, const double *B, double *C) {
    double tmp[12];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            tmp[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
        tmp[9 + i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
    }
    memcpy(C, tmp, sizeof tmp);
}
static void se3_identity(double *T) { memset(T, 0, 12 * sizeof(double)); T[0] = T[4] = T[8] = 1.0; }

static void joint_motion(int type, const double *ax, double q, double *T) {
    se3_identity(T);
    if (type == JT_REVOLUTE) {
        double c = cos(q), s = sin(q), C = 1.0 - c, x = ax[0], y = ax[1], z = ax[2];
        T[0] = c + x * x * C;     T[1] = x * y * C - z * s; T[2] = x * z * C + y * s;
        T[3] = y * x * C + z * s; T[4] = c + y * y * C;     T[5] = y * z * C - x * s;
        T[6] = z * x * C - y * s; T[7] = z * y * C + x * s; T[8] = c + z * z * C;
    } else if (type == JT_PRISMATIC) {
        T[9] = ax[0] * q; T[10] = ax[1] * q; T[11] = ax[2] * q;
    }
}

/* Forward kinematics + geometry placement.  oMi [njoints][12], oMg [ngeoms][12]. */
void oracle_fk(const oracle_model *m, const double *q, double *oMi, double *oMg) {
    se3_identity(oMi);
    for (int j = 1; j < m->njoints; j++) {
        double Tj[12], tmp[12];
        joint_motion(m->joint_type[j], m->joint_axis + 3 * j, q[m->joint_idx_q[j]], Tj);
        se3_mul(oMi + 12 * m->joint_parent[j], m->joint_placement + 12 * j, tmp);
        se3_mul(tmp, Tj, oMi + 12 * j);
    }
    for (int g = 0; g < m->ngeoms; g++)
        se3_mul(oMi + 12 * m->geom_parent[g], m->geom_placement + 12 * g, oMg + 12 * g);
}

/* ------------------------------------------------------------------ support functions */
typedef struct {
    int shape;
    const double *par;
    const double *verts;
    int nverts;
    const double *T; /* world placement */
    double radius;   /* sphere / capsule inflation handled outside GJK */
} shape_t;

static void support_local(const shape_t *s, const double *d, double *out, oracle_counters *cnt) {
    cnt->support_calls++;
    switch (s->shape) {
    case SH_SPHERE:
        out[0] = out[1] = out[2] = 0.0;
        break;
    case SH_CAPSULE:
        out[0] = out[1] = 0.0;
        out[2] = d[2] >= 0.0 ? s->par[1] : -s->par[1];
        break;
    case SH_BOX:
        for (int i = 0; i < 3; i++) out[i] = d[i] >= 0.0 ? s->par[i] : -s->par[i];
        break;
    case SH_CYLINDER: {
        double n = sqrt(d[0] * d[0] + d[1] * d[1]);
        if (n > 0.0) { out[0] = s->par[0] * d[0] / n; out[1] = s->par[0] * d[1] / n; }
        else { out[0] = out[1] = 0.0; }
        out[2] = d[2] >= 0.0 ? s->par[1] : -s->par[1];
        break;
    }
    default: {
        int best = 0;
        double bd = -INFINITY;
        for (int i = 0; i < s->nverts; i++) {
            double v = dot3(s->verts + 3 * i, d);
            if (v > bd) { bd = v; best = i; }
        }
        cnt->support_verts += s->nverts;
        memcpy(out, s->verts + 3 * best, 3 * sizeof(double));
    }
    }
}

static void support_world(const shape_t *s, const double *d, double *out, oracle_counters *cnt) {
    const double *R = s->T;
    double dl[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
                    R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
    double p[3];
    support_local(s, dl, p, cnt);
    for (int i = 0; i < 3; i++) out[i] = R[3 * i] * p[0] + R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2] + R[9 + i];
}

/* ------------------------------------------------------------------ simplex closest point */
typedef struct {
    double w[4][3], a[4][3], b[4][3];
    double lam[4];
    int n;
} simplex_t;

static void simplex_keep(simplex_t *s, const int *idx, const double *lam, int n) {
    simplex_t t = *s;
    for (int i = 0; i < n; i++) {
        memcpy(s->w[i], t.w[idx[i]], sizeof s->w[i]);
        memcpy(s->a[i], t.a[idx[i]], sizeof s->a[i]);
        memcpy(s->b[i], t.b[idx[i]], sizeof s->b[i]);
        s->lam[i] = lam[i];
    }
    s->n = n;
}

/* closest point to the origin on segment (A,B): barycentric (u,v) */
static void closest_segment(const double *A, const double *B, double *lam, int *mask) {
    double ab[3];
    sub3(B, A, ab);
    double den = dot3(ab, ab);
    double t = den > 0.0 ? -dot3(A, ab) / den : 0.0;
    if (t <= 0.0) { lam[0] = 1.0; lam[1] = 0.0; *mask = 1; }
    else if (t >= 1.0) { lam[0] = 0.0; lam[1] = 1.0; *mask = 2; }
    else { lam[0] = 1.0 - t; lam[1] = t; *mask = 3; }
}

/* closest point to the origin on triangle (A,B,C) (Voronoi-region walk); mask = used verts */
static void closest_triangle(const double *A, const double *B, const double *C, double *lam, int *mask) {
    double ab[3], ac[3], ap[3], bp[3], cp[3];
    sub3(B, A, ab); sub3(C, A, ac);
    for (int i = 0; i < 3; i++) { ap[i] = -A[i]; bp[i] = -B[i]; cp[i] = -C[i]; }
    double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; *mask = 1; return; }
    double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; *mask = 2; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        double v = d1 / (d1 - d3);
        lam[0] = 1 - v; lam[1] = v; lam[2] = 0; *mask = 3; return;
    }
    double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; *mask = 4; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        double w = d2 / (d2 - d6);
        lam[0] = 1 - w; lam[1] = 0; lam[2] = w; *mask = 5; return;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        lam[0] = 0; lam[1] = 1 - w; lam[2] = w; *mask = 6; return;
    }
    double den = 1.0 / (va + vb + vc);
    lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1.0 - lam[1] - lam[2]; *mask = 7;
}

/* Reduce the simplex to the face closest to the origin; v = closest point.
 * Returns 1 when the origin is inside a tetrahedron (shapes intersect). */
static int simplex_closest(simplex_t *s, double *v) {
    double lam[4];
    int idx[4], n = 0;
    if (s->n == 1) {
        lam[0] = 1.0; idx[0] = 0; n = 1;
    } else if (s->n == 2) {
        int mask; double l2[2];
        closest_segment(s->w[0], s->w[1], l2, &mask);
        for (int i = 0; i < 2; i++) if (mask & (1 << i)) { idx[n] = i; lam[n++] = l2[i]; }
    } else if (s->n == 3) {
        int mask; double l3[3];
        closest_triangle(s->w[0], s->w[1], s->w[2], l3, &mask);
        for (int i = 0; i < 3; i++) if (mask & (1 << i)) { idx[n] = i; lam[n++] = l3[i]; }
    } else {
        /* tetrahedron: test the origin against each face's outward side */
        static const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
        double best = INFINITY, bl[3] = {0, 0, 0};
        int bmask = 0, bf = -1, outside_any = 0;
        for (int f = 0; f < 4; f++) {
            const double *A = s->w[F[f][0]], *B = s->w[F[f][1]], *C = s->w[F[f][2]], *D = s->w[F[f][3]];
            double ab[3], ac[3], nrm[3], ad[3];
            sub3(B, A, ab); sub3(C, A, ac); cross3(ab, ac, nrm); sub3(D, A, ad);
            double sd = dot3(nrm, ad);      /* side of the opposite vertex */
            double so = -dot3(nrm, A);      /* side of the origin */
            int flat = sd * sd <= 1e-24 * dot3(nrm, nrm) * dot3(ad, ad); /* degenerate: no reliable inside */
            int outside = flat ? 1 : ((sd > 0.0) ? (so <= 0.0) : (so >= 0.0));
            if (!outside) continue;
            outside_any = 1;
            double l3[3]; int mask;
            closest_triangle(A, B, C, l3, &mask);
            double p[3];
            for (int i = 0; i < 3; i++) p[i] = l3[0] * A[i] + l3[1] * B[i] + l3[2] * C[i];
            double d2 = dot3(p, p);
            if (d2 < best) { best = d2; bf = f; bmask = mask; memcpy(bl, l3, sizeof bl); }
        }
        if (!outside_any) {
            /* barycentric coordinates of the origin in the tetrahedron (for the certificate) */
            double vol[4], tot = 0.0;
            for (int f = 0; f < 4; f++) {
                const double *A = s->w[F[f][0]], *B = s->w[F[f][1]], *C = s->w[F[f][2]];
                double bc[3];
                cross3(B, C, bc);
                vol[F[f][3]] = fabs(dot3(A, bc));
                tot += vol[F[f][3]];
            }
            for (int i = 0; i < 4; i++) s->lam[i] = tot > 0.0 ? vol[i] / tot : 0.25;
            v[0] = v[1] = v[2] = 0.0;
            return 1;
        }
        for (int i = 0; i < 3; i++) if (bmask & (1 << i)) { idx[n] = F[bf][i]; lam[n++] = bl[i]; }
    }
    simplex_keep(s, idx, lam, n);
    for (int k = 0; k < 3; k++) {
        v[k] = 0.0;
        for (int i = 0; i < s->n; i++) v[k] += s->lam[i] * s->w[i][k];
    }
    return 0;
}

#define ORACLE_GJK_MAX_ITERS 2000
#define ORACLE_GJK_TOL 1e-11

/* Distance between the *cores* of two shapes; returns 0 and intersect=1 when cores overlap.
 * pa/pb: witness points on core A / core B (equal when intersecting). */
static double gjk_core_distance(const shape_t *A, const shape_t *B, double *pa, double *pb, int *intersect,
                                oracle_counters *cnt) {
    simplex_t s;
    s.n = 0;
    double v[3] = {A->T[9] - B->T[9], A->T[10] - B->T[10], A->T[11] - B->T[11]};
    if (dot3(v, v) == 0.0) { v[0] = 1.0; }
    *intersect = 0;
    double best_lb = 0.0;
    for (int it = 0; it < ORACLE_GJK_MAX_ITERS; it++) {
        cnt->gjk_iters++;
        double d[3] = {-v[0], -v[1], -v[2]}, sa[3], sb[3], w[3];
        support_world(A, d, sa, cnt);
        support_world(B, v, sb, cnt);
        sub3(sa, sb, w);
        double vv = dot3(v, v), vw = dot3(v, w);
        if (s.n > 0) {
            if (vv - vw <= ORACLE_GJK_TOL * sqrt(vv)) break; /* ub - lb <= tol */
            int dup = 0;
            for (int i = 0; i < s.n; i++)
                if (s.w[i][0] == w[0] && s.w[i][1] == w[1] && s.w[i][2] == w[2]) dup = 1;
            if (dup) break;
        }
        /* <v,w>/|v| is a lower bound of the distance for ANY v: once it is positive the shapes are
         * certainly apart.  Curved supports (cylinders) produce sliver tetrahedra near convergence
         * whose inside test is rounding noise; a claimed enclosure that contradicts the bound is
         * discarded and the previous (converged to within ub - lb) closest point is kept. */
        if (vw > 0.0 && vw / sqrt(vv) > best_lb) best_lb = vw / sqrt(vv);
        simplex_t prev = s;
        double v_prev[3] = {v[0], v[1], v[2]};
        memcpy(s.w[s.n], w, sizeof w); memcpy(s.a[s.n], sa, sizeof sa); memcpy(s.b[s.n], sb, sizeof sb);
        s.n++;
        const int enclosed = simplex_closest(&s, v);
        if ((enclosed || dot3(v, v) <= 1e-28) && best_lb > 1e-9) { s = prev; memcpy(v, v_prev, sizeof v_prev); break; }
        if (enclosed) { *intersect = 1; break; }
        if (dot3(v, v) <= 1e-28) { *intersect = 1; break; }
        if (dot3(v, v) > vv && s.n > 1 && best_lb > 1e-9 && vv - best_lb * sqrt(vv) < 1e-9 * (1.0 + vv)) { s = prev; memcpy(v, v_prev, sizeof v_prev); break; }
    }
    for (int k = 0; k < 3; k++) {
        pa[k] = pb[k] = 0.0;
        for (int i = 0; i < s.n; i++) { pa[k] += s.lam[i] * s.a[i][k]; pb[k] += s.lam[i] * s.b[i][k]; }
    }
    if (*intersect) return 0.0;
    return sqrt(dot3(v, v));
}

static void make_shape(const oracle_model *m, int g, const double *oMg, shape_t *s) {
    s->shape = m->geom_shape[g];
    s->par = m->geom_params + 4 * g;
    s->verts = m->verts + 3 * m->geom_vert_off[g];
    s->nverts = m->geom_vert_cnt[g];
    s->T = oMg + 12 * g;
    s->radius = (s->shape == SH_SPHERE || s->shape == SH_CAPSULE) ? s->par[0] : 0.0;
}

/* Distance between geometries ga and gb (>= 0; 0 when they touch or overlap).
 * out[0..2] witness on A, out[3..5] witness on B (world frame), out[6] = 1 if intersecting. */
double oracle_pair_distance(const oracle_model *m, const double *oMg, int ga, int gb, double *out,
                            oracle_counters *cnt) {
    shape_t A, B;
    oracle_counters local = {0};
    if (!cnt) cnt = &local;
    make_shape(m, ga, oMg, &A);
    make_shape(m, gb, oMg, &B);
    cnt->pair_tests++;
    double pa[3], pb[3];
    int inter;
    double dc = gjk_core_distance(&A, &B, pa, pb, &inter, cnt);
    double r = A.radius + B.radius;
    double dist;
    if (inter || dc <= r) {
        dist = 0.0;
        inter = 1;
        if (dc > 0.0 && !isnan(dc)) {
            /* cores apart but inflated shapes overlap: any point of the overlap works as witness */
            double t = A.radius / (r > 0.0 ? r : 1.0); /* |p-pa| = dc*rA/r <= rA, |p-pb| = dc*rB/r <= rB */
            for (int k = 0; k < 3; k++) { double c = pa[k] + (pb[k] - pa[k]) * t; pa[k] = pb[k] = c; }
        }
    } else {
        dist = dc - r;
        double n[3] = {(pb[0] - pa[0]) / dc, (pb[1] - pa[1]) / dc, (pb[2] - pa[2]) / dc};
        for (int k = 0; k < 3; k++) { pa[k] += A.radius * n[k]; pb[k] -= B.radius * n[k]; }
    }
    if (out) {
        memcpy(out, pa, 3 * sizeof(double)); memcpy(out + 3, pb, 3 * sizeof(double));
        out[6] = inter ? 1.0 : 0.0;
    }
    return dist;
}

static int sphere_cull(const oracle_model *m, const double *oMg, int ga, int gb, double margin) {
    double ca[3], cb[3];
    const double *Ta = oMg + 12 * ga, *Tb = oMg + 12 * gb, *sa = m->geom_bsphere + 4 * ga, *sb = m->geom_bsphere + 4 * gb;
    for (int i = 0; i < 3; i++) {
        ca[i] = Ta[3 * i] * sa[0] + Ta[3 * i + 1] * sa[1] + Ta[3 * i + 2] * sa[2] + Ta[9 + i];
        cb[i] = Tb[3 * i] * sb[0] + Tb[3 * i + 1] * sb[1] + Tb[3 * i + 2] * sb[2] + Tb[9 + i];
    }
    double d[3];
    sub3(ca, cb, d);
    double r = sa[3] + sb[3] + margin + 1e-9;
    return dot3(d, d) > r * r;
}

static double point_box_dist(const double *T, const double *h, const double *p) {
    /* distance from world point p to the box with placement T and half sides h */
    double d2 = 0.0;
    for (int i = 0; i < 3; i++) {
        double l = T[i] * (p[0] - T[9]) + T[3 + i] * (p[1] - T[10]) + T[6 + i] * (p[2] - T[11]);
        double e = fabs(l) - h[i];
        if (e > 0.0) d2 += e * e;
    }
    return sqrt(d2);
}

static void sphere_world(const double *T, const double *s, double *c) {
    for (int i = 0; i < 3; i++) c[i] = T[3 * i] * s[0] + T[3 * i + 1] * s[1] + T[3 * i + 2] * s[2] + T[9 + i];
}

/* Conservative proxy bounds of one pair (enclosing / inscribed spheres, exact boxes): the same
 * broadphase idea the CUDA engine uses, in float64.  lb <= distance <= ub (ub may be < 0). */
static void proxy_bounds(const oracle_model *m, const double *oMg, int ga, int gb, double *lb, double *ub) {
    int box = m->geom_shape[ga] == SH_BOX ? ga : (m->geom_shape[gb] == SH_BOX ? gb : -1);
    if (box >= 0) {
        int ot = box == ga ? gb : ga;
        double c[3], ic[3];
        sphere_world(oMg + 12 * ot, m->geom_bsphere + 4 * ot, c);
        sphere_world(oMg + 12 * ot, m->geom_isphere + 4 * ot, ic);
        *lb = point_box_dist(oMg + 12 * box, m->geom_params + 4 * box, c) - m->geom_bsphere[4 * ot + 3];
        *ub = point_box_dist(oMg + 12 * box, m->geom_params + 4 * box, ic) - m->geom_isphere[4 * ot + 3];
    } else {
        double ca[3], cb[3], ia[3], ib[3], d[3], e[3];
        sphere_world(oMg + 12 * ga, m->geom_bsphere + 4 * ga, ca);
        sphere_world(oMg + 12 * gb, m->geom_bsphere + 4 * gb, cb);
        sphere_world(oMg + 12 * ga, m->geom_isphere + 4 * ga, ia);
        sphere_world(oMg + 12 * gb, m->geom_isphere + 4 * gb, ib);
        sub3(ca, cb, d); sub3(ia, ib, e);
        *lb = sqrt(dot3(d, d)) - m->geom_bsphere[4 * ga + 3] - m->geom_bsphere[4 * gb + 3];
        *ub = sqrt(dot3(e, e)) - m->geom_isphere[4 * ga + 3] - m->geom_isphere[4 * gb + 3];
    }
}

/* One configuration.  Returns hit (0/1).  Restates computeCollisions(..., stop_at_first) +
 * np.any(isCollision) (reference core/utils.py:48-51).  When want_all != 0 every pair is
 * evaluated (as computeDistances does, core/utils.py:86) and *min_dist is the min over pairs.
 * first_pair = index of the first colliding pair in list order, or -1.
 * use_cull: skip pairs whose bounding spheres are farther apart than `padding` (cannot change
 * the verdict; min_dist then is min over the un-culled pairs and the cull lower bounds). */
double oracle_soup_pair_distance(const oracle_model *m, const double *oMg, int ga, int gb, const double *triA, int ntA,
                                 const double *triB, int ntB, double stop_at, int mode);

/* Distance of a pair under the model's semantics.  Solids: the hull / primitive distance.  With
 * surface triangles: a surface lies inside its hull, so the hull distance is a lower bound and only
 * a pair whose hulls are within `padding` needs the triangle test; its result (stopped at the first
 * triangle pair within padding) replaces the hull distance. */
static double soup_pair_distance_from(const oracle_model *m, const double *oMg, int ga, int gb, const double *triA, int ntA,
                                      const double *triB, int ntB, double stop_at, int mode, double best0);

static int pair_has_surface(const oracle_model *m, int ga, int gb) { return m->tris && (m->tri_cnt[ga] > 0 || m->tri_cnt[gb] > 0); }

/* *is_exact (optional): 1 when the value returned is the distance under the model's semantics, 0 when
 * it is the hull distance of a pair with a surface (a lower bound of its surface distance). */
static double pair_distance_semantics(const oracle_model *m, const double *oMg, int ga, int gb, double padding, int want_exact,
                                      int *is_exact, oracle_counters *cnt) {
    double d = oracle_pair_distance(m, oMg, ga, gb, NULL, cnt);
    if (is_exact) *is_exact = !pair_has_surface(m, ga, gb);
    if (d <= padding && pair_has_surface(m, ga, gb)) {
        d = oracle_soup_pair_distance(m, oMg, ga, gb, m->tris + 9 * (int64_t)m->tri_off[ga], m->tri_cnt[ga],
                                      m->tris + 9 * (int64_t)m->tri_off[gb], m->tri_cnt[gb], want_exact ? -1.0 : padding, want_exact ? 0 : 2);
        if (is_exact) *is_exact = 1;
    }
    return d;
}

/* How far a near-convex mesh surface lies below its hull (Panda: 0.39 mm, pyroboplan_b200.engine.mesh_dent_depth):
 * only the size of the FIRST, cheap search window of the exact triangle distance -- not an assumption. */
#define ORACLE_MAX_DENT 1e-3

int oracle_check_state(const oracle_model *m, const double *q, double padding, int want_all, int use_cull,
                       double *min_dist, int32_t *first_pair, double *oMi_scratch, double *oMg_scratch,
                       oracle_counters *cnt) {
    oracle_fk(m, q, oMi_scratch, oMg_scratch);
    int hit = 0;
    double md = INFINITY;
    int fp = -1;
    if (use_cull == 2 && !want_all) {
        /* Work-counting mode (verdict unchanged): proxy bounds first -- any certain hit decides
         * the state without narrowphase -- then GJK only on the undecided pairs.  Used to count
         * the algorithmic narrowphase work of a cull-aware method (F_alg). */
        const double slack = 1e-9;
        for (int k = 0; k < m->npairs; k++) {
            double lb, ub;
            proxy_bounds(m, oMg_scratch, m->pairs[2 * k], m->pairs[2 * k + 1], &lb, &ub);
            if (m->tris && (m->tri_cnt[m->pairs[2 * k]] > 0 || m->tri_cnt[m->pairs[2 * k + 1]] > 0)) continue;   /* overlapping inscribed spheres do not prove a SURFACE contact */
            if (ub <= padding - slack) { if (first_pair) *first_pair = k; if (min_dist) *min_dist = 0.0; return 1; }
        }
        for (int k = 0; k < m->npairs; k++) {
            double lb, ub;
            int ga = m->pairs[2 * k], gb = m->pairs[2 * k + 1];
            proxy_bounds(m, oMg_scratch, ga, gb, &lb, &ub);
            if (lb > padding + slack) { cnt->culled++; continue; }
            double d = pair_distance_semantics(m, oMg_scratch, ga, gb, padding, 0, NULL, cnt);
            if (d < md) md = d;
            if (d <= padding) { hit = 1; fp = k; break; }
        }
        if (min_dist) *min_dist = md;
        if (first_pair) *first_pair = fp;
        return hit;
    }
    /* With surfaces and a wanted minimum distance: a hull distance is only a lower bound of the
     * pair's surface distance (by at most the dents of both meshes), so the pairs that could define
     * the minimum are re-evaluated with the exact triangle distance afterwards. */
    const int refine = min_dist != NULL && m->tris != NULL && want_all;
    double *dk = NULL;
    unsigned char *exact_k = NULL;
    if (refine) {
        dk = malloc(sizeof(double) * (size_t)(m->npairs > 0 ? m->npairs : 1));
        exact_k = calloc((size_t)(m->npairs > 0 ? m->npairs : 1), 1);
        for (int k = 0; k < m->npairs; k++) dk[k] = INFINITY;
    }
    for (int k = 0; k < m->npairs; k++) {
        int ga = m->pairs[2 * k], gb = m->pairs[2 * k + 1];
        if (use_cull && !want_all && sphere_cull(m, oMg_scratch, ga, gb, padding)) { cnt->culled++; continue; }
        /* (with surfaces the running minimum is a minimum of HULL distances, a lower bound only: no cull on it) */
        if (use_cull && want_all && !refine && md < INFINITY && sphere_cull(m, oMg_scratch, ga, gb, md)) { cnt->culled++; continue; }
        int is_exact = 1;
        double d = pair_distance_semantics(m, oMg_scratch, ga, gb, padding, min_dist != NULL, &is_exact, cnt);
        if (refine) { dk[k] = d; exact_k[k] = (unsigned char)is_exact; }
        if (d < md) md = d;
        if (d <= padding) {
            if (fp < 0) fp = k;
            hit = 1;
            if (!want_all) break;
        }
    }
    if (refine) {
        for (;;) {
            int kb = -1;
            for (int k = 0; k < m->npairs; k++) if (kb < 0 || dk[k] < dk[kb]) kb = k;
            if (kb < 0 || exact_k[kb] || !isfinite(dk[kb])) { if (kb >= 0) md = dk[kb]; break; }
            const int ga = m->pairs[2 * kb], gb = m->pairs[2 * kb + 1];
            /* cheap first: within a couple of dent depths of the hull distance (always enough for near-convex
             * meshes, the Panda's); a non-convex mesh (UR upper arm, forearm: decimetres below the hull) falls
             * through to the unbounded search */
            const double first_try = dk[kb] + 2.0 * ORACLE_MAX_DENT;
            double ds = soup_pair_distance_from(m, oMg_scratch, ga, gb, m->tris + 9 * (int64_t)m->tri_off[ga], m->tri_cnt[ga],
                                                m->tris + 9 * (int64_t)m->tri_off[gb], m->tri_cnt[gb], -1.0, 0, first_try);
            if (!(ds < first_try))
                ds = soup_pair_distance_from(m, oMg_scratch, ga, gb, m->tris + 9 * (int64_t)m->tri_off[ga], m->tri_cnt[ga],
                                             m->tris + 9 * (int64_t)m->tri_off[gb], m->tri_cnt[gb], -1.0, 0, INFINITY);
            dk[kb] = ds;
            exact_k[kb] = 1;
        }
        free(dk); free(exact_k);
    }
    if (min_dist) *min_dist = md;
    if (first_pair) *first_pair = fp;
    return hit;
}

/* ------------------------------------------------------------------ pthread parallel-for */
typedef void (*chunk_fn)(void *ctx, int64_t begin, int64_t end, oracle_counters *cnt, int64_t *aux);
typedef struct {
    chunk_fn fn; void *ctx; int64_t n, chunk; atomic_llong *next;
    oracle_counters cnt; int64_t aux;
} worker_t;

static void *worker_main(void *arg) {
    worker_t *w = (worker_t *)arg;
    for (;;) {
        int64_t b = atomic_fetch_add(w->next, w->chunk);
        if (b >= w->n) break;
        int64_t e = b + w->chunk < w->n ? b + w->chunk : w->n;
        w->fn(w->ctx, b, e, &w->cnt, &w->aux);
    }
    return NULL;
}

int oracle_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(int64_t n, int64_t chunk, int nthreads, chunk_fn fn, void *ctx, oracle_counters *total,
                         int64_t *aux_total) {
    if (nthreads <= 0) nthreads = oracle_num_threads();
    if (nthreads > 256) nthreads = 256;
    atomic_llong next = 0;
    worker_t w[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; t++) {
        memset(&w[t], 0, sizeof w[t]);
        w[t].fn = fn; w[t].ctx = ctx; w[t].n = n; w[t].chunk = chunk; w[t].next = &next;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, worker_main, &w[t]);
    worker_main(&w[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    oracle_counters sum = {0};
    int64_t aux = 0;
    for (int t = 0; t < nthreads; t++) {
        sum.pair_tests += w[t].cnt.pair_tests; sum.gjk_iters += w[t].cnt.gjk_iters;
        sum.support_calls += w[t].cnt.support_calls; sum.support_verts += w[t].cnt.support_verts;
        sum.culled += w[t].cnt.culled; aux += w[t].aux;
    }
    if (total) *total = sum;
    if (aux_total) *aux_total = aux;
}

typedef struct {
    const oracle_model *m; const double *q; double padding; int want_all, use_cull;
    uint8_t *hit; double *min_dist; int32_t *first_pair;
} states_ctx;

static void states_chunk(void *vctx, int64_t b, int64_t e, oracle_counters *cnt, int64_t *aux) {
    states_ctx *c = (states_ctx *)vctx;
    const oracle_model *m = c->m;
    (void)aux;
    double *oMi = malloc(sizeof(double) * 12 * m->njoints), *oMg = malloc(sizeof(double) * 12 * m->ngeoms);
    for (int64_t i = b; i < e; i++) {
        double md;
        int32_t fp;
        c->hit[i] = (uint8_t)oracle_check_state(m, c->q + i * m->nq, c->padding, c->want_all, c->use_cull, &md, &fp,
                                                oMi, oMg, cnt);
        if (c->min_dist) c->min_dist[i] = md;
        if (c->first_pair) c->first_pair[i] = fp;
    }
    free(oMi); free(oMg);
}

/* Batch of configurations, nthreads host threads (<= 0: all online cores). */
void oracle_check_states(const oracle_model *m, const double *q, int64_t n, double padding, int want_all,
                         int use_cull, int nthreads, uint8_t *hit, double *min_dist, int32_t *first_pair,
                         oracle_counters *total) {
    states_ctx c = {m, q, padding, want_all, use_cull, hit, min_dist, first_pair};
    parallel_for(n, 256, nthreads, states_chunk, &c, total, NULL);
}

/* Number of samples of one discretised segment: ceil(||dq|| / step) + 1
 * (reference planning/utils.py:137). */
int64_t oracle_segment_samples(const double *q1, const double *q2, int nq, double step) {
    double s = 0.0;
    for (int i = 0; i < nq; i++) { double d = q2[i] - q1[i]; s += d * d; }
    return (int64_t)ceil(sqrt(s) / step) + 1;
}

/* Write the samples of one segment: q1 + linspace(0,1,n)[i] * (q2 - q1)
 * (reference planning/utils.py:138-139; numpy.linspace: i * (1/(n-1)), last sample = 1.0). */
void oracle_discretize_segment(const double *q1, const double *q2, int nq, int64_t n, double *out) {
    double step = n > 1 ? 1.0 / (double)(n - 1) : 0.0;
    for (int64_t i = 0; i < n; i++) {
        double s = (n > 1 && i == n - 1) ? 1.0 : (double)i * step;
        for (int k = 0; k < nq; k++) out[i * nq + k] = q1[k] + s * (q2[k] - q1[k]);
    }
}

typedef struct {
    const oracle_model *m; const double *q1, *q2; double step, padding; int use_cull;
    uint8_t *hit; int32_t *first_bad;
} edges_ctx;

static void edges_chunk(void *vctx, int64_t b0, int64_t e0, oracle_counters *cnt, int64_t *evals) {
    edges_ctx *c = (edges_ctx *)vctx;
    const oracle_model *m = c->m;
    double *oMi = malloc(sizeof(double) * 12 * m->njoints), *oMg = malloc(sizeof(double) * 12 * m->ngeoms);
    double *qs = malloc(sizeof(double) * m->nq);
    for (int64_t e = b0; e < e0; e++) {
        const double *a = c->q1 + e * m->nq, *b = c->q2 + e * m->nq;
        int h = 0;
        int32_t fb = -1;
        (*evals)++;
        if (oracle_check_state(m, b, c->padding, 0, c->use_cull, NULL, NULL, oMi, oMg, cnt)) {
            h = 1;
            fb = -2; /* destination itself collides */
        } else {
            int64_t n = oracle_segment_samples(a, b, m->nq, c->step);
            double st = n > 1 ? 1.0 / (double)(n - 1) : 0.0;
            for (int64_t i = 0; i < n; i++) {
                double s = (n > 1 && i == n - 1) ? 1.0 : (double)i * st;
                for (int k = 0; k < m->nq; k++) qs[k] = a[k] + s * (b[k] - a[k]);
                (*evals)++;
                if (oracle_check_state(m, qs, c->padding, 0, c->use_cull, NULL, NULL, oMi, oMg, cnt)) {
                    h = 1; fb = (int32_t)i; break;
                }
            }
        }
        c->hit[e] = (uint8_t)h;
        if (c->first_bad) c->first_bad[e] = fb;
    }
    free(oMi); free(oMg); free(qs);
}

/* Edge validity as has_collision_free_path computes it (reference planning/utils.py:89-111):
 * q2 checked first, then the discretised path sequentially, stopping at the first colliding
 * state (core/utils.py:125-130).  hit[e] = 1 means the edge is NOT collision free.
 * states_evaluated counts the states the sequential reference would have evaluated;
 * first_bad[e] = index of the first colliding sample along the path (-2: q2 itself), -1 if none. */
void oracle_check_edges(const oracle_model *m, const double *q1, const double *q2, int64_t n_edges, double step,
                        double padding, int use_cull, int nthreads, uint8_t *hit, int32_t *first_bad,
                        int64_t *states_evaluated, oracle_counters *total) {
    edges_ctx c = {m, q1, q2, step, padding, use_cull, hit, first_bad};
    parallel_for(n_edges, 16, nthreads, edges_chunk, &c, total, states_evaluated);
}

/* Pre-discretised path, as check_collisions_along_path takes it (reference core/utils.py:90-132). */
int oracle_check_path(const oracle_model *m, const double *q_path, int64_t n, double padding, int use_cull) {
    double *oMi = malloc(sizeof(double) * 12 * m->njoints), *oMg = malloc(sizeof(double) * 12 * m->ngeoms);
    oracle_counters c = {0};
    int h = 0;
    for (int64_t i = 0; i < n && !h; i++)
        h = oracle_check_state(m, q_path + i * m->nq, padding, 0, use_cull, NULL, NULL, oMi, oMg, &c);
    free(oMi); free(oMg);
    return h;
}

/* ------------------------------------------------------------------ penetration depth (EPA)
 * coal reports a NEGATIVE min_distance for overlapping shapes (the penetration depth, found with
 * the Expanding Polytope Algorithm, Van den Bergen 2001), the nearest points as the deepest points
 * and the normal from shape 1 to shape 2.  The reference reads exactly those fields
 * (/root/reference/src/pyroboplan/core/utils.py:505-516;
 * /root/reference/src/pyroboplan/ik/nullspace_components.py:130-141).  Restated here in float64 on
 * the same support functions: the GJK tetrahedron that encloses the origin of A - B is expanded
 * face by face until the face closest to the origin is a face of A - B itself (to 1e-10). */
#define EPA_MAXV 320
#define EPA_MAXF 1024
#define EPA_TOL 1e-10

typedef struct { int v[3]; double n[3]; double d; int alive; } epa_face_t;

static int epa_make_face(epa_face_t *f, int i0, int i1, int i2, double (*W)[3]) {
    double e1[3], e2[3], n[3];
    sub3(W[i1], W[i0], e1); sub3(W[i2], W[i0], e2);
    cross3(e1, e2, n);
    double len = sqrt(dot3(n, n));
    f->v[0] = i0; f->v[1] = i1; f->v[2] = i2; f->alive = 1;
    if (len < 1e-300) { f->n[0] = f->n[1] = f->n[2] = 0.0; f->d = INFINITY; return 0; }
    for (int k = 0; k < 3; k++) f->n[k] = n[k] / len;
    f->d = dot3(f->n, W[i0]);
    if (f->d < 0.0) {   /* orient away from the origin (which is inside) */
        f->v[1] = i2; f->v[2] = i1;
        for (int k = 0; k < 3; k++) f->n[k] = -f->n[k];
        f->d = -f->d;
    }
    return 1;
}

/* depth >= 0 of the cores' overlap; n = unit normal from A to B; pa/pb deepest points on the cores.
 * Returns 0 when no polytope around the origin can be started from the simplex handed over (a single
 * point: the cores touch in a vertex). */
static int epa_core(const shape_t *A, const shape_t *B, const simplex_t *s, double *depth, double *n_out, double *pa, double *pb,
                    oracle_counters *cnt) {
    static __thread double W[EPA_MAXV][3], Av[EPA_MAXV][3];
    static __thread epa_face_t F[EPA_MAXF];
    static __thread int edges[3 * EPA_MAXF][2];
    int nv = 0, nf = 0;
    if (s->n < 2 || s->n > 4) return 0;
    for (int i = 0; i < s->n; i++) { memcpy(W[i], s->w[i], sizeof W[i]); memcpy(Av[i], s->a[i], sizeof Av[i]); }
    nv = s->n;
    if (s->n == 4) {
        static const int tet[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
        /* the GJK orientation test uses the opposite vertex; here the origin: it is inside */
        for (int f = 0; f < 4; f++) epa_make_face(&F[nf++], tet[f][0], tet[f][1], tet[f][2], W);
    } else {
        /* GJK stopped on a triangle / segment THROUGH the origin (it does whenever the support
         * points stay in one plane by symmetry, e.g. a sphere centre inside a cylinder: every
         * direction lies in the axial plane through the centre).  That is an overlap, not a
         * touching contact: close the simplex into a bipyramid with extra support points -- two
         * apexes along +- the triangle normal, or a ring of three around the segment. */
        double dirs[3][3];
        int nd = 0;
        if (s->n == 3) {
            double e1[3], e2[3], nn[3];
            sub3(W[1], W[0], e1); sub3(W[2], W[0], e2); cross3(e1, e2, nn);
            const double len = sqrt(dot3(nn, nn));
            if (!(len > 1e-150)) return 0;
            for (int k = 0; k < 3; k++) { dirs[0][k] = nn[k] / len; dirs[1][k] = -nn[k] / len; }
            nd = 2;
        } else {
            double u[3], e[3] = {0, 0, 0}, d1[3], d2[3];
            sub3(W[1], W[0], u);
            if (!(dot3(u, u) > 1e-300)) return 0;
            const double ax = fabs(u[0]), ay = fabs(u[1]), az = fabs(u[2]);
            e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
            cross3(u, e, d1); cross3(u, d1, d2);
            const double l1 = sqrt(dot3(d1, d1)), l2 = sqrt(dot3(d2, d2));
            for (int k = 0; k < 3; k++) {
                d1[k] /= l1; d2[k] /= l2;
                dirs[0][k] = d1[k];
                dirs[1][k] = -0.5 * d1[k] + 0.86602540378443865 * d2[k];
                dirs[2][k] = -0.5 * d1[k] - 0.86602540378443865 * d2[k];
            }
            nd = 3;
        }
        for (int j = 0; j < nd; j++) {
            double md[3] = {-dirs[j][0], -dirs[j][1], -dirs[j][2]}, sa[3], sb[3];
            support_world(A, dirs[j], sa, cnt);
            support_world(B, md, sb, cnt);
            sub3(sa, sb, W[nv]); memcpy(Av[nv], sa, sizeof sa);
            nv++;
        }
        if (s->n == 3) {
            static const int bip[6][3] = {{0, 1, 3}, {1, 2, 3}, {2, 0, 3}, {0, 1, 4}, {1, 2, 4}, {2, 0, 4}};
            for (int f = 0; f < 6; f++) epa_make_face(&F[nf++], bip[f][0], bip[f][1], bip[f][2], W);
        } else {
            static const int bip[6][3] = {{2, 3, 0}, {3, 4, 0}, {4, 2, 0}, {2, 3, 1}, {3, 4, 1}, {4, 2, 1}};
            for (int f = 0; f < 6; f++) epa_make_face(&F[nf++], bip[f][0], bip[f][1], bip[f][2], W);
        }
    }
    int best = -1;
    for (int iter = 0; iter < 4 * EPA_MAXV; iter++) {
        best = -1;
        for (int f = 0; f < nf; f++)
            if (F[f].alive && (best < 0 || F[f].d < F[best].d)) best = f;
        if (best < 0 || !isfinite(F[best].d)) return 0;
        double d[3] = {F[best].n[0], F[best].n[1], F[best].n[2]}, md[3] = {-d[0], -d[1], -d[2]}, sa[3], sb[3], w[3];
        support_world(A, d, sa, cnt);
        support_world(B, md, sb, cnt);
        sub3(sa, sb, w);
        const double proj = dot3(w, d);
        if (proj - F[best].d <= EPA_TOL * (1.0 + F[best].d) || nv >= EPA_MAXV) break;
        memcpy(W[nv], w, sizeof w); memcpy(Av[nv], sa, sizeof sa);
        /* remove the faces that see the new point, collect their edges */
        int ne = 0;
        for (int f = 0; f < nf; f++) {
            if (!F[f].alive) continue;
            double rel[3];
            sub3(w, W[F[f].v[0]], rel);
            if (isfinite(F[f].d) ? dot3(F[f].n, rel) > 1e-14 : 1) {
                F[f].alive = 0;
                for (int k = 0; k < 3; k++) { edges[ne][0] = F[f].v[k]; edges[ne][1] = F[f].v[(k + 1) % 3]; ne++; }
            }
        }
        if (ne == 0) break;   /* numerically on the surface already */
        /* horizon = edges whose reverse is not among the removed faces' edges */
        for (int e = 0; e < ne; e++) {
            int shared = 0;
            for (int g = 0; g < ne && !shared; g++)
                if (g != e && edges[g][0] == edges[e][1] && edges[g][1] == edges[e][0]) shared = 1;
            if (shared) continue;
            int slot = -1;
            for (int f = 0; f < nf && slot < 0; f++) if (!F[f].alive) slot = f;
            if (slot < 0) { if (nf >= EPA_MAXF) return 0; slot = nf++; }
            epa_make_face(&F[slot], edges[e][0], edges[e][1], nv, W);
        }
        nv++;
    }
    if (best < 0) return 0;
    /* Several coplanar triangles can share the minimal plane distance (a polygonal face of A - B):
     * take the one that contains the projection of the origin, so that the barycentric weights are
     * a convex combination (the witnesses then lie in their shapes). */
    const double dmin = F[best].d;
    double bu = 0.0, bv = 0.0, bviol = INFINITY;
    int bf = best;
    for (int fi = 0; fi < nf; fi++) {
        const epa_face_t *f = &F[fi];
        if (!f->alive || !(f->d <= dmin + 1e-9 * (1.0 + dmin))) continue;
        double mm[3] = {f->n[0] * f->d, f->n[1] * f->d, f->n[2] * f->d};
        double e1[3], e2[3], r[3];
        sub3(W[f->v[1]], W[f->v[0]], e1); sub3(W[f->v[2]], W[f->v[0]], e2); sub3(mm, W[f->v[0]], r);
        const double a11 = dot3(e1, e1), a12 = dot3(e1, e2), a22 = dot3(e2, e2), b1 = dot3(r, e1), b2 = dot3(r, e2);
        const double det = a11 * a22 - a12 * a12;
        if (!(det > 0.0)) continue;
        const double u = (b1 * a22 - b2 * a12) / det, v = (b2 * a11 - b1 * a12) / det;
        const double viol = fmax(fmax(-u, -v), u + v - 1.0);
        if (viol < bviol) { bviol = viol; bu = u; bv = v; bf = fi; }
    }
    const epa_face_t *f = &F[bf];
    *depth = f->d;
    memcpy(n_out, f->n, 3 * sizeof(double));
    if (bu < 0.0) bu = 0.0;
    if (bv < 0.0) bv = 0.0;
    if (bu + bv > 1.0) { const double t = bu + bv; bu /= t; bv /= t; }
    double m[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        pa[k] = Av[f->v[0]][k] + bu * (Av[f->v[1]][k] - Av[f->v[0]][k]) + bv * (Av[f->v[2]][k] - Av[f->v[0]][k]);
        m[k] = W[f->v[0]][k] + bu * (W[f->v[1]][k] - W[f->v[0]][k]) + bv * (W[f->v[2]][k] - W[f->v[0]][k]);
        pb[k] = pa[k] - m[k];
    }
    return 1;
}

/* GJK that keeps its final simplex (the same loop as gjk_core_distance) */
static double gjk_core_distance_simplex(const shape_t *A, const shape_t *B, double *pa, double *pb, int *intersect, simplex_t *sout,
                                        double *vout, oracle_counters *cnt) {
    simplex_t s;
    s.n = 0;
    double v[3] = {A->T[9] - B->T[9], A->T[10] - B->T[10], A->T[11] - B->T[11]};
    if (dot3(v, v) == 0.0) { v[0] = 1.0; }
    *intersect = 0;
    double best_lb = 0.0;
    for (int it = 0; it < ORACLE_GJK_MAX_ITERS; it++) {
        cnt->gjk_iters++;
        double d[3] = {-v[0], -v[1], -v[2]}, sa[3], sb[3], w[3];
        support_world(A, d, sa, cnt);
        support_world(B, v, sb, cnt);
        sub3(sa, sb, w);
        double vv = dot3(v, v), vw = dot3(v, w);
        if (s.n > 0) {
            if (vv - vw <= ORACLE_GJK_TOL * sqrt(vv)) break;
            int dup = 0;
            for (int i = 0; i < s.n; i++)
                if (s.w[i][0] == w[0] && s.w[i][1] == w[1] && s.w[i][2] == w[2]) dup = 1;
            if (dup) break;
        }
        if (vw > 0.0 && vw / sqrt(vv) > best_lb) best_lb = vw / sqrt(vv);   /* see gjk_core_distance */
        simplex_t prev = s;
        double v_prev[3] = {v[0], v[1], v[2]};
        memcpy(s.w[s.n], w, sizeof w); memcpy(s.a[s.n], sa, sizeof sa); memcpy(s.b[s.n], sb, sizeof sb);
        s.n++;
        simplex_t before = s;
        const int enclosed = simplex_closest(&s, v);
        if ((enclosed || dot3(v, v) <= 1e-28) && best_lb > 1e-9) { s = prev; memcpy(v, v_prev, sizeof v_prev); break; }
        if (enclosed) { *intersect = 1; s = before; break; }   /* keep the enclosing tetrahedron */
        if (dot3(v, v) <= 1e-28) { *intersect = 1; break; }
    }
    for (int k = 0; k < 3; k++) {
        pa[k] = pb[k] = 0.0;
        if (!*intersect || s.n < 4)
            for (int i = 0; i < s.n; i++) { pa[k] += s.lam[i] * s.a[i][k]; pb[k] += s.lam[i] * s.b[i][k]; }
    }
    *sout = s;
    memcpy(vout, v, 3 * sizeof(double));
    if (*intersect) return 0.0;
    return sqrt(dot3(v, v));
}

/* Signed distance between geometries ga and gb: > 0 separated, < 0 penetration depth (coal's
 * convention).  out[0..2] nearest / deepest point on A, out[3..5] on B, out[6..8] unit normal from A
 * to B, out[9] = 1 when the result is exact in the sense above, 0 when the shapes touch in a way
 * the polytope expansion cannot start from (then the distance returned is 0). */
double oracle_pair_signed_distance(const oracle_model *m, const double *oMg, int ga, int gb, double *out, oracle_counters *cnt) {
    shape_t A, B;
    oracle_counters local = {0};
    if (!cnt) cnt = &local;
    make_shape(m, ga, oMg, &A);
    make_shape(m, gb, oMg, &B);
    double pa[3], pb[3], v[3], n[3] = {0, 0, 0};
    int inter, ok = 1;
    simplex_t s;
    double dc = gjk_core_distance_simplex(&A, &B, pa, pb, &inter, &s, v, cnt);
    const double r = A.radius + B.radius;
    double dist;
    if (!inter) {
        for (int k = 0; k < 3; k++) n[k] = (pb[k] - pa[k]) / dc;
        dist = dc - r;   /* negative when only the inflated shapes overlap */
    } else {
        double depth = 0.0;
        if (epa_core(&A, &B, &s, &depth, n, pa, pb, cnt)) dist = -depth - r;
        else { dist = -r; ok = (s.n == 1 && r > 0.0) ? 1 : 0; }   /* cores touch in a vertex: the depth is that of the radii */
        if (!ok || (dist == -r && depth == 0.0 && dot3(n, n) == 0.0)) {
            double len = sqrt(dot3(v, v));
            if (len > 0.0) for (int k = 0; k < 3; k++) n[k] = -v[k] / len;
        }
    }
    for (int k = 0; k < 3; k++) { pa[k] += A.radius * n[k]; pb[k] -= B.radius * n[k]; }
    if (out) {
        memcpy(out, pa, 3 * sizeof(double)); memcpy(out + 3, pb, 3 * sizeof(double)); memcpy(out + 6, n, 3 * sizeof(double));
        out[9] = ok ? 1.0 : 0.0;
    }
    return dist;
}

/* ------------------------------------------------------------------ triangle-soup semantics (study)
 * Pinocchio loads a URDF <mesh> as a coal BVHModel: the SURFACE triangles, not the solid and not
 * its convex hull (Appendix A.5 of SURVEY.md; the call sites are
 * /root/reference/src/pyroboplan/models/panda.py:27 -> pinocchio.buildModelsFromUrdf).  Two meshes
 * collide when two of their triangles are within the security margin; a primitive collides with a
 * mesh when it is within the margin of one triangle.  The engine works on convex hulls; for the
 * Panda meshes the hull and the surface differ only by sub-millimetre dents.  The
 * functions below evaluate the surface semantics exactly (float64, brute force over triangles with
 * bounding-box rejection) so that the difference can be COUNTED on the benchmark inputs. */
static inline double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }

/* squared distance between segments p1q1 and p2q2 (Ericson, Real-Time Collision Detection 5.1.9) */
static double seg_seg_dist2(const double *p1, const double *q1, const double *p2, const double *q2) {
    double d1[3], d2[3], r[3];
    sub3(q1, p1, d1); sub3(q2, p2, d2); sub3(p1, p2, r);
    const double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    double s, t;
    if (a <= 0.0 && e <= 0.0) return dot3(r, r);
    if (a <= 0.0) { s = 0.0; t = clamp01(f / e); }
    else {
        const double c = dot3(d1, r);
        if (e <= 0.0) { t = 0.0; s = clamp01(-c / a); }
        else {
            const double b = dot3(d1, d2), den = a * e - b * b;
            s = den > 0.0 ? clamp01((b * f - c * e) / den) : 0.0;
            t = (b * s + f) / e;
            if (t < 0.0) { t = 0.0; s = clamp01(-c / a); }
            else if (t > 1.0) { t = 1.0; s = clamp01((b - c) / a); }
        }
    }
    double d[3];
    for (int k = 0; k < 3; k++) d[k] = (p1[k] + s * d1[k]) - (p2[k] + t * d2[k]);
    return dot3(d, d);
}

/* squared distance from point p to triangle abc (Ericson 5.1.5, Voronoi regions) */
static double point_tri_dist2(const double *p, const double *a, const double *b, const double *c) {
    double ab[3], ac[3], ap[3], bp[3], cp[3], q[3];
    sub3(b, a, ab); sub3(c, a, ac); sub3(p, a, ap);
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) return dot3(ap, ap);
    sub3(p, b, bp);
    const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) return dot3(bp, bp);
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
        const double v = d1 / (d1 - d3);
        for (int k = 0; k < 3; k++) q[k] = ap[k] - v * ab[k];
        return dot3(q, q);
    }
    sub3(p, c, cp);
    const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) return dot3(cp, cp);
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
        const double w = d2 / (d2 - d6);
        for (int k = 0; k < 3; k++) q[k] = ap[k] - w * ac[k];
        return dot3(q, q);
    }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        for (int k = 0; k < 3; k++) q[k] = bp[k] - w * (c[k] - b[k]);
        return dot3(q, q);
    }
    const double den = 1.0 / (va + vb + vc), v = vb * den, w = vc * den;
    for (int k = 0; k < 3; k++) q[k] = ap[k] - v * ab[k] - w * ac[k];
    return dot3(q, q);
}

/* does the segment pq cross the triangle abc (closed sets, non-coplanar case; the coplanar case is
 * caught by the edge-edge distances, which are then 0) */
static int seg_crosses_tri(const double *p, const double *q, const double *a, const double *b, const double *c) {
    double ab[3], ac[3], n[3], ap[3], aq[3];
    sub3(b, a, ab); sub3(c, a, ac); cross3(ab, ac, n);
    sub3(p, a, ap); sub3(q, a, aq);
    const double sp = dot3(n, ap), sq = dot3(n, aq);
    if ((sp > 0.0 && sq > 0.0) || (sp < 0.0 && sq < 0.0) || sp == sq) return 0;
    const double t = sp / (sp - sq);
    double x[3], ax[3];
    for (int k = 0; k < 3; k++) x[k] = p[k] + t * (q[k] - p[k]);
    sub3(x, a, ax);
    const double d00 = dot3(ab, ab), d01 = dot3(ab, ac), d11 = dot3(ac, ac), d20 = dot3(ax, ab), d21 = dot3(ax, ac);
    const double den = d00 * d11 - d01 * d01;
    if (!(den > 0.0)) return 0;
    const double v = (d11 * d20 - d01 * d21) / den, w = (d00 * d21 - d01 * d20) / den;
    return v >= 0.0 && w >= 0.0 && v + w <= 1.0;
}

/* exact distance between two triangles (0 when they intersect); A = a[0..8], B = b[0..8] */
static double tri_tri_dist2(const double *a, const double *b) {
    for (int i = 0; i < 3; i++) {
        if (seg_crosses_tri(a + 3 * i, a + 3 * ((i + 1) % 3), b, b + 3, b + 6)) return 0.0;
        if (seg_crosses_tri(b + 3 * i, b + 3 * ((i + 1) % 3), a, a + 3, a + 6)) return 0.0;
    }
    double best = INFINITY;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double d = seg_seg_dist2(a + 3 * i, a + 3 * ((i + 1) % 3), b + 3 * j, b + 3 * ((j + 1) % 3));
            if (d < best) best = d;
        }
    for (int i = 0; i < 3; i++) {
        double d = point_tri_dist2(a + 3 * i, b, b + 3, b + 6);
        if (d < best) best = d;
        d = point_tri_dist2(b + 3 * i, a, a + 3, a + 6);
        if (d < best) best = d;
    }
    return best;
}

/* exported for the tests: distance between two triangles given as 9 doubles each */
double oracle_tri_tri_distance(const double *a, const double *b) { return sqrt(tri_tri_dist2(a, b)); }

static void tri_to_world(const double *T, const double *tri, double *out, double *lo, double *hi) {
    for (int k = 0; k < 3; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    for (int v = 0; v < 3; v++)
        for (int k = 0; k < 3; k++) {
            const double x = T[3 * k] * tri[3 * v] + T[3 * k + 1] * tri[3 * v + 1] + T[3 * k + 2] * tri[3 * v + 2] + T[9 + k];
            out[3 * v + k] = x;
            if (x < lo[k]) lo[k] = x;
            if (x > hi[k]) hi[k] = x;
        }
}

/* Distance between the SURFACES of geometries ga and gb, where a geometry with ntri > 0 is its
 * triangle soup (tri: [ntri][9] vertex coordinates in the geometry frame) and a geometry with
 * ntri == 0 is the solid primitive of the model.  Stops as soon as a distance <= stop_at (or 0) is
 * found and returns it.  mode 0: exact minimum otherwise, with bounding-box rejection; mode 1: the
 * same by brute force (the tests validate mode 0 with it); mode 2: verdict only -- triangles
 * farther than stop_at from the other mesh's box are never visited, and when nothing is within
 * stop_at the value returned is merely some number > stop_at. */
static double soup_pair_distance_from(const oracle_model *m, const double *oMg, int ga, int gb, const double *triA, int ntA,
                                      const double *triB, int ntB, double stop_at, int mode, double best0);

double oracle_soup_pair_distance(const oracle_model *m, const double *oMg, int ga, int gb, const double *triA, int ntA,
                                 const double *triB, int ntB, double stop_at, int mode) {
    return soup_pair_distance_from(m, oMg, ga, gb, triA, ntA, triB, ntB, stop_at, mode, INFINITY);
}

/* best0: a known upper bound of the result (INFINITY if none) -- triangle pairs farther apart than
 * it are rejected by their boxes.  If nothing is closer than best0, best0 is returned. */
static double soup_pair_distance_from(const oracle_model *m, const double *oMg, int ga, int gb, const double *triA, int ntA,
                                      const double *triB, int ntB, double stop_at, int mode, double best0) {
    oracle_counters cnt = {0};
    double best = best0;
    const int brute = mode == 1;
    if (stop_at < 0.0) stop_at = 0.0;
    if (ntA > 0 && ntB > 0) {
        double *WB = malloc((size_t)ntB * 15 * sizeof(double));
        for (int j = 0; j < ntB; j++) tri_to_world(oMg + 12 * gb, triB + 9 * j, WB + 15 * j, WB + 15 * j + 9, WB + 15 * j + 12);
        double blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int j = 0; j < ntB; j++)
            for (int k = 0; k < 3; k++) {
                if (WB[15 * j + 9 + k] < blo[k]) blo[k] = WB[15 * j + 9 + k];
                if (WB[15 * j + 12 + k] > bhi[k]) bhi[k] = WB[15 * j + 12 + k];
            }
        /* the triangle pairs that can matter lie in the overlap of the two bounding boxes (grown by
         * the distance that is still interesting): keep only B's triangles that reach A's box */
        int nB = ntB;
        if (!brute) {
            double alo[3] = {INFINITY, INFINITY, INFINITY}, ahi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = 0; i < ntA; i++) {
                double wa[9], lo[3], hi[3];
                tri_to_world(oMg + 12 * ga, triA + 9 * i, wa, lo, hi);
                for (int k = 0; k < 3; k++) { if (lo[k] < alo[k]) alo[k] = lo[k]; if (hi[k] > ahi[k]) ahi[k] = hi[k]; }
            }
            /* an upper bound of the surface distance: the boxes' diameter sum is crude but finite */
            const double reach = stop_at > 0.0 ? stop_at : 0.0;
            int any_overlap = 1;
            for (int k = 0; k < 3; k++) if (alo[k] - bhi[k] > reach || blo[k] - ahi[k] > reach) any_overlap = 0;
            if (!any_overlap && mode == 2) { free(WB); return INFINITY; }
            if (any_overlap && mode == 2) {
                /* verdict mode: only triangles within `reach` of the other box can produce a hit; if
                 * none does, the exact distance is still wanted by the caller only as "> stop_at" */
                nB = 0;
                for (int j = 0; j < ntB; j++) {
                    const double *wb = WB + 15 * j;
                    int keep = 1;
                    for (int k = 0; k < 3; k++) if (wb[9 + k] - ahi[k] > reach || alo[k] - wb[12 + k] > reach) keep = 0;
                    if (keep) { if (nB != j) memcpy(WB + 15 * nB, wb, 15 * sizeof(double)); nB++; }
                }
                if (nB == 0) { free(WB); return INFINITY; }
                for (int k = 0; k < 3; k++) { blo[k] = INFINITY; bhi[k] = -INFINITY; }
                for (int j = 0; j < nB; j++)
                    for (int k = 0; k < 3; k++) {
                        if (WB[15 * j + 9 + k] < blo[k]) blo[k] = WB[15 * j + 9 + k];
                        if (WB[15 * j + 12 + k] > bhi[k]) bhi[k] = WB[15 * j + 12 + k];
                    }
                best = best0;
            }
        }
        const int verdict_only = mode == 2;
        const double reach2 = stop_at > 0.0 ? stop_at * stop_at : 0.0;
        for (int i = 0; i < ntA && !(best <= stop_at); i++) {
            double wa[9], lo[3], hi[3];
            tri_to_world(oMg + 12 * ga, triA + 9 * i, wa, lo, hi);
            if (!brute) {   /* against B's whole box first */
                double g2 = 0.0;
                for (int k = 0; k < 3; k++) { const double g = fmax(fmax(lo[k] - bhi[k], blo[k] - hi[k]), 0.0); g2 += g * g; }
                if (g2 >= best * best || (verdict_only && g2 > reach2)) continue;
            }
            for (int j = 0; j < nB; j++) {
                const double *wb = WB + 15 * j;
                if (!brute) {
                    double g2 = 0.0;
                    for (int k = 0; k < 3; k++) { const double g = fmax(fmax(lo[k] - wb[12 + k], wb[9 + k] - hi[k]), 0.0); g2 += g * g; }
                    if (g2 >= best * best || (verdict_only && g2 > reach2)) continue;
                }
                const double d2 = tri_tri_dist2(wa, wb);
                if (d2 < best * best) best = sqrt(d2);
                if (best <= stop_at) break;
            }
        }
        free(WB);
        return best;
    }
    if (ntA == 0 && ntB == 0) return oracle_pair_distance(m, oMg, ga, gb, NULL, NULL);
    /* primitive (solid) against a soup: GJK between the primitive's core and each triangle */
    const int gp = ntA == 0 ? ga : gb, gs = ntA == 0 ? gb : ga;
    const double *tri = ntA == 0 ? triB : triA;
    const int nt = ntA == 0 ? ntB : ntA;
    shape_t P, Tr;
    make_shape(m, gp, oMg, &P);
    Tr.shape = SH_CONVEX; Tr.par = NULL; Tr.nverts = 3; Tr.T = oMg + 12 * gs; Tr.radius = 0.0;
    double pc[3];
    {
        const double *Tp = oMg + 12 * gp, *bs = m->geom_bsphere + 4 * gp;
        for (int k = 0; k < 3; k++) pc[k] = Tp[3 * k] * bs[0] + Tp[3 * k + 1] * bs[1] + Tp[3 * k + 2] * bs[2] + Tp[9 + k];
    }
    const double pr = m->geom_bsphere[4 * gp + 3];
    for (int i = 0; i < nt && !(best <= stop_at); i++) {
        Tr.verts = tri + 9 * i;
        if (mode != 1) {   /* the triangle's box against the primitive's bounding sphere */
            double wt[9], lo[3], hi[3], g2 = 0.0;
            tri_to_world(oMg + 12 * gs, tri + 9 * i, wt, lo, hi);
            for (int k = 0; k < 3; k++) { const double g = fmax(fmax(lo[k] - pc[k], pc[k] - hi[k]), 0.0); g2 += g * g; }
            const double lb = sqrt(g2) - pr;
            if (lb >= best || (mode == 2 && lb > stop_at)) continue;
        }
        double pa[3], pb[3];
        int inter;
        double d = gjk_core_distance(&P, &Tr, pa, pb, &inter, &cnt);
        d = inter ? 0.0 : fmax(d - P.radius, 0.0);
        if (d < best) best = d;
    }
    return best;
}

/* Batch: per state the hull verdict (the engine's semantics: any active pair with hull distance <=
 * padding) and the surface verdict (any active pair with surface distance <= padding).  A surface
 * lies inside its hull, so only pairs that hit by hull need the triangle test.  min_hull_hit[i] =
 * number of pairs of state i that hit by hull but not by surface.  tris: all triangles [..][9] in
 * their geometry frames; tri_off / tri_cnt per geometry (cnt 0 = primitive). */
typedef struct {
    const oracle_model *m; const double *q; double padding;
    const double *tris; const int32_t *tri_off, *tri_cnt;
    uint8_t *hull_hit, *soup_hit; int32_t *flipped_pairs;
} soup_ctx;

static void soup_chunk(void *vctx, int64_t b, int64_t e, oracle_counters *cnt, int64_t *aux) {
    soup_ctx *c = (soup_ctx *)vctx;
    const oracle_model *m = c->m;
    double *oMi = malloc(sizeof(double) * 12 * m->njoints), *oMg = malloc(sizeof(double) * 12 * m->ngeoms);
    for (int64_t i = b; i < e; i++) {
        oracle_fk(m, c->q + i * m->nq, oMi, oMg);
        int hh = 0, sh = 0, flipped = 0;
        for (int p = 0; p < m->npairs; p++) {
            const int ga = m->pairs[2 * p], gb = m->pairs[2 * p + 1];
            double ca[3], cb[3];
            for (int k = 0; k < 3; k++) {
                const double *Ta = oMg + 12 * ga, *Tb = oMg + 12 * gb, *sa = m->geom_bsphere + 4 * ga, *sb = m->geom_bsphere + 4 * gb;
                ca[k] = Ta[3 * k] * sa[0] + Ta[3 * k + 1] * sa[1] + Ta[3 * k + 2] * sa[2] + Ta[9 + k];
                cb[k] = Tb[3 * k] * sb[0] + Tb[3 * k + 1] * sb[1] + Tb[3 * k + 2] * sb[2] + Tb[9 + k];
            }
            double dcen[3];
            sub3(ca, cb, dcen);
            const double rr = m->geom_bsphere[4 * ga + 3] + m->geom_bsphere[4 * gb + 3] + c->padding;
            if (dot3(dcen, dcen) > rr * rr) continue;
            if (oracle_pair_distance(m, oMg, ga, gb, NULL, cnt) > c->padding) continue;
            hh = 1;
            const double ds = oracle_soup_pair_distance(m, oMg, ga, gb, c->tris + 9 * (int64_t)c->tri_off[ga], c->tri_cnt[ga],
                                                        c->tris + 9 * (int64_t)c->tri_off[gb], c->tri_cnt[gb], c->padding, 2);
            (*aux)++;
            if (ds <= c->padding) sh = 1; else flipped++;
        }
        c->hull_hit[i] = (uint8_t)hh; c->soup_hit[i] = (uint8_t)sh;
        if (c->flipped_pairs) c->flipped_pairs[i] = flipped;
    }
    free(oMi); free(oMg);
}

void oracle_soup_check_states(const oracle_model *m, const double *q, int64_t n, double padding, const double *tris,
                              const int32_t *tri_off, const int32_t *tri_cnt, int nthreads, uint8_t *hull_hit, uint8_t *soup_hit,
                              int32_t *flipped_pairs, int64_t *soup_tests) {
    soup_ctx c = {m, q, padding, tris, tri_off, tri_cnt, hull_hit, soup_hit, flipped_pairs};
    parallel_for(n, 64, nthreads, soup_chunk, &c, NULL, soup_tests);
}
