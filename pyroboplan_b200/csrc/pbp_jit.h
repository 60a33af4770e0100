// Model-specialised broadphase: CUDA source generated from the compiled model, built with NVRTC
// for the engine's device at creation time.
//
// The generic k_broadphase walks the joint / geometry / pair TABLES in shared memory: every pair
// test pays for loading its pair record, its thresholds and the proxy centres (thread-interleaved
// shared memory), ~60 instructions per pair and ~7800 per state for the Panda.  Here the model is
// the program: the joint chain is straight-line code with the placements as literals (zeros and
// ones fold away), the proxy centres of the moving geometries stay in registers, every pair test
// is unrolled with its geometry ids resolved at generation time, and the per-call thresholds sit
// in the kernel parameter (constant bank) -- ~14 instructions per pair.  The item append works on
// each thread's own items with shared-memory atomics.  Only the boolean mode (the planning hot path: states and
// edge samples) is specialised; distance / first-pair modes and models that do not fit use the
// generic kernel.  Set PBP_NO_JIT=1 to disable.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pbp_model.cuh"

namespace pbp {
namespace jit {

constexpr int SPEC_THREADS = 128;
constexpr int SPEC_MAX_PAIRS = 300;   // thresholds travel in the kernel parameter (4 KB limit)

// Kernel parameter of the generated kernel (mirrored textually in the generated source).
struct SpecArgs {
    const float *q, *q1, *q2;
    const int32_t *group;
    const float *frac;
    int32_t mode;
    int32_t _pad;
    long long n_states;
    unsigned char *hit;
    int32_t *bin_state;
    uint32_t *bin_count;
    long long cap;
    float thr[SPEC_MAX_PAIRS][3];   // free / hit / capsule-free squared thresholds; only the first npairs entries are passed
};
inline size_t spec_args_bytes(int npairs) { return offsetof(SpecArgs, thr) + (size_t)npairs * 12; }

// ---- tiny expression helper: a scalar that is either a literal or a named variable
struct Ex {
    bool lit;
    float v;
    std::string name;
    static Ex L(float x) { return Ex{true, x, ""}; }
    static Ex V(const std::string &n) { return Ex{false, 0.f, n}; }
    bool zero() const { return lit && v == 0.f; }
    std::string str() const {
        if (!lit) return name;
        char b[48];
        snprintf(b, sizeof b, "%.9gf", (double)v);
        std::string s(b);
        if (s.find('.') == std::string::npos && s.find('e') == std::string::npos && s.find("inf") == std::string::npos)
            s.insert(s.size() - 1, ".0");
        return "(" + s + ")";
    }
};

struct Gen {
    std::string out;
    int tmp = 0;
    void line(const std::string &s) { out += "    " + s + "\n"; }
    // emits `const float NAME = sum_i a[i]*b[i] (+ add);` folding literals; returns the value
    Ex dot(const std::string &name, const std::vector<Ex> &a, const std::vector<Ex> &b, const Ex *add = nullptr) {
        float acc = add && add->lit ? add->v : 0.f;
        std::vector<std::pair<Ex, Ex>> terms;   // (coefficient, variable) with at least one non-literal
        for (size_t i = 0; i < a.size(); i++) {
            if (a[i].zero() || b[i].zero()) continue;
            if (a[i].lit && b[i].lit) { acc += a[i].v * b[i].v; continue; }
            terms.push_back({a[i], b[i]});
        }
        std::string tail;
        bool have_tail = false;
        if (add && !add->lit) { tail = add->name; have_tail = true; }
        if (acc != 0.f) {
            const std::string lit = Ex::L(acc).str();
            tail = have_tail ? "(" + tail + " + " + lit + ")" : lit;
            have_tail = true;
        }
        if (terms.empty()) {
            if (!have_tail) return Ex::L(0.f);
            if (!(add && !add->lit)) return Ex::L(acc);
            line("const float " + name + " = " + tail + ";");
            return Ex::V(name);
        }
        // fmaf chain, innermost = last term (+ tail)
        std::string e;
        for (size_t i = terms.size(); i-- > 0;) {
            const Ex &c = terms[i].first, &x = terms[i].second;
            std::string prod_a = c.str(), prod_b = x.str();
            const bool unit = (c.lit && fabsf(c.v) == 1.f) || (x.lit && fabsf(x.v) == 1.f);
            if (e.empty() && !have_tail) {
                if (unit) {
                    const Ex &other = (c.lit && fabsf(c.v) == 1.f) ? x : c;
                    const float sgn = (c.lit && fabsf(c.v) == 1.f) ? c.v : x.v;
                    e = sgn < 0 ? "(-" + other.str() + ")" : other.str();
                } else e = "(" + prod_a + " * " + prod_b + ")";
            } else {
                const std::string inner = e.empty() ? tail : e;
                e = "fmaf(" + prod_a + ", " + prod_b + ", " + inner + ")";
            }
        }
        line("const float " + name + " = " + e + ";");
        return Ex::V(name);
    }
};

struct Xf { Ex m[12]; };   // rotation row-major [0..8], translation [9..11]

inline Xf identity_xf() {
    Xf x;
    for (int i = 0; i < 12; i++) x.m[i] = Ex::L((i == 0 || i == 4 || i == 8) ? 1.f : 0.f);
    return x;
}

// C = A * B with B literal (a joint placement)
inline Xf mul_lit(Gen &g, const std::string &pfx, const Xf &A, const float *B) {
    Xf C;
    for (int i = 0; i < 3; i++) {
        const std::vector<Ex> row = {A.m[3 * i], A.m[3 * i + 1], A.m[3 * i + 2]};
        for (int c = 0; c < 3; c++)
            C.m[3 * i + c] = g.dot(pfx + "_" + std::to_string(3 * i + c), row, {Ex::L(B[c]), Ex::L(B[3 + c]), Ex::L(B[6 + c])});
        C.m[9 + i] = g.dot(pfx + "_" + std::to_string(9 + i), row, {Ex::L(B[9]), Ex::L(B[10]), Ex::L(B[11])}, &A.m[9 + i]);
    }
    return C;
}

inline Ex act_lit(Gen &g, const std::string &name, const Xf &X, int row, const float *p) {
    return g.dot(name, {X.m[3 * row], X.m[3 * row + 1], X.m[3 * row + 2]}, {Ex::L(p[0]), Ex::L(p[1]), Ex::L(p[2])}, &X.m[9 + row]);
}

struct Input {
    int nq, njoints, ngeoms, npairs;
    const DevJoint *joints;
    const BpGeom *bpgeoms;     // [ngeoms]
    const int32_t *geom_parent; // [ngeoms]
    const BpBox *bpboxes;
    const BpPair *bppairs;     // internal order
    const BpGroup *groups;
    int ngroups;
};

inline std::string generate_broadphase(const Input &in) {
    Gen g;
    const int NQ = in.nq, NP = in.npairs, NW = (NP + 31) / 32;
    std::string s;
    s += "// generated by pbp_jit.h -- model-specialised boolean broadphase\n";
    s += "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t;\n";
    s += "struct SpecArgs { const float *q, *q1, *q2; const int32_t *group; const float *frac; int32_t mode; int32_t _pad;\n"
         "  long long n_states; unsigned char *hit; int32_t *bin_state; uint32_t *bin_count; long long cap; float thr[" +
         std::to_string(NP) + "][3]; };\n";
    s += "#define NQ " + std::to_string(NQ) + "\n#define NP " + std::to_string(NP) + "\n#define NW " + std::to_string(NW) +
         "\n#define BT " + std::to_string(SPEC_THREADS) + "\n";
    // squared distance point <-> segment (p0, u = p1 - p0, 1 / |u|^2) and segment <-> segment (Ericson, Real-Time Collision Detection 5.1.9)
    s += "__device__ __forceinline__ float pt_seg2(float cx, float cy, float cz, float px, float py, float pz, float ux, float uy, float uz, float inv_uu) {\n"
         "    const float dx = cx - px, dy = cy - py, dz = cz - pz;\n"
         "    const float t = fminf(fmaxf(fmaf(dx, ux, fmaf(dy, uy, dz * uz)) * inv_uu, 0.f), 1.f);\n"
         "    const float wx = fmaf(-t, ux, dx), wy = fmaf(-t, uy, dy), wz = fmaf(-t, uz, dz);\n"
         "    return fmaf(wx, wx, fmaf(wy, wy, wz * wz));\n}\n";
    s += "__device__ __forceinline__ float seg_seg2(float px, float py, float pz, float d1x, float d1y, float d1z, float a,\n"
         "                                          float qx, float qy, float qz, float d2x, float d2y, float d2z, float e) {\n"
         "    const float rx = px - qx, ry = py - qy, rz = pz - qz;\n"
         "    const float f = fmaf(d2x, rx, fmaf(d2y, ry, d2z * rz)), c = fmaf(d1x, rx, fmaf(d1y, ry, d1z * rz));\n"
         "    const float b = fmaf(d1x, d2x, fmaf(d1y, d2y, d1z * d2z));\n"
         "    const float den = a * e - b * b;\n"
         "    float sN = den > 1e-12f * a * e ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;\n"
         "    float tN = (b * sN + f) / e;\n"
         "    if (tN < 0.f) { tN = 0.f; sN = fminf(fmaxf(-c / a, 0.f), 1.f); }\n"
         "    else if (tN > 1.f) { tN = 1.f; sN = fminf(fmaxf((b - c) / a, 0.f), 1.f); }\n"
         "    const float wx = fmaf(sN, d1x, rx) - tN * d2x, wy = fmaf(sN, d1y, ry) - tN * d2y, wz = fmaf(sN, d1z, rz) - tN * d2z;\n"
         "    return fmaf(wx, wx, fmaf(wy, wy, wz * wz));\n}\n";
    s += "extern \"C\" __global__ void __launch_bounds__(BT) k_bp_spec(const SpecArgs A) {\n";
    s += "    __shared__ __align__(16) float sq[BT * NQ];\n    __shared__ uint32_t scnt[NP];\n    __shared__ uint32_t sbase[NP];\n";
    s += "    const int tid = threadIdx.x;\n    const int64_t block_first = (int64_t)blockIdx.x * BT;\n";
    s += "    const int count = (int)min((int64_t)BT, A.n_states - block_first);\n";
    // state load (as load_states of the generic kernel)
    s += "    if (A.mode == 0) {\n"
         "        const float *base = A.q + block_first * NQ;\n        const int total = count * NQ;\n"
         "        if ((((unsigned long long)base) & 15ull) == 0ull) {\n"
         "            const int n4 = total >> 2;\n            const float4 *b4 = reinterpret_cast<const float4 *>(base);\n"
         "            float4 *s4 = reinterpret_cast<float4 *>(sq);\n"
         "            for (int i = tid; i < n4; i += BT) s4[i] = __ldg(b4 + i);\n"
         "            for (int i = (n4 << 2) + tid; i < total; i += BT) sq[i] = __ldg(base + i);\n"
         "        } else {\n            for (int i = tid; i < total; i += BT) sq[i] = __ldg(base + i);\n        }\n"
         "    } else if (tid < count) {\n"
         "        const int64_t st = block_first + tid;\n        const int64_t gg = A.group[st];\n        const float f = A.frac[st];\n"
         "        const float *a = A.q1 + gg * NQ, *b = A.q2 + gg * NQ;\n"
         "        for (int k = 0; k < NQ; k++) { const float qa = __ldg(a + k), qb = __ldg(b + k); sq[tid * NQ + k] = fmaf(f, qb - qa, qa); }\n"
         "    }\n";
    s += "    for (int i = tid; i < NP; i += BT) scnt[i] = 0u;\n    __syncthreads();\n";
    s += "    const int64_t state = block_first + tid;\n    const bool valid = tid < count;\n";
    s += "    const int32_t group = valid ? (A.group ? A.group[state] : (int32_t)state) : 0;\n";
    s += "    bool alive = valid;\n    if (A.group && valid && A.hit[group]) alive = false;\n    bool state_hit = false;\n";
    for (int w = 0; w < NW; w++) s += "    uint32_t w" + std::to_string(w) + " = 0u;\n";
    s += "    if (__any_sync(0xffffffffu, alive)) {\n";
    g.out.clear();
    // ---- joint values
    for (int k = 0; k < NQ; k++) g.line("const float q" + std::to_string(k) + " = valid ? sq[tid * NQ + " + std::to_string(k) + "] : 0.f;");
    // ---- forward kinematics, straight-line
    std::vector<Xf> X(in.njoints);
    X[0] = identity_xf();
    for (int j = 1; j < in.njoints; j++) {
        const DevJoint &J = in.joints[j];
        const std::string pj = "j" + std::to_string(j);
        const Xf Am = mul_lit(g, pj + "a", X[J.parent], J.P);
        const std::string qn = "q" + std::to_string(J.idx_q);
        Xf R = Am;
        if (J.type == PBP_JOINT_REVOLUTE) {
            g.line("float " + pj + "s, " + pj + "c; sincosf(" + qn + ", &" + pj + "s, &" + pj + "c);");
            const Ex sn = Ex::V(pj + "s"), cs = Ex::V(pj + "c");
            if (J.axis_code != 0) {
                // rotation about +-x / y / z: two columns mix (same convention as rotate_columns of pbp_model.cuh)
                const int ax = J.axis_code - 1;
                const int U = ax == 0 ? 1 : 0, W = ax == 2 ? 1 : 2;
                const float sgn = J.axis[ax] * (ax == 1 ? -1.f : 1.f);   // Ry mixes its columns with the opposite sign
                g.line("const float " + pj + "sg = " + (sgn < 0 ? std::string("-") : std::string("")) + pj + "s;");
                const Ex sg = Ex::V(pj + "sg");
                g.line("const float " + pj + "ng = -" + pj + "sg;");
                const Ex ng = Ex::V(pj + "ng");
                for (int i = 0; i < 3; i++) {
                    const Ex au = Am.m[3 * i + U], aw = Am.m[3 * i + W];
                    R.m[3 * i + U] = g.dot(pj + "r" + std::to_string(3 * i + U), {au, aw}, {cs, sg});
                    R.m[3 * i + W] = g.dot(pj + "r" + std::to_string(3 * i + W), {aw, au}, {cs, ng});
                }
            } else {
                // generic axis: Rodrigues matrix, then a full product
                const float x = J.axis[0], y = J.axis[1], z = J.axis[2];
                g.line("const float " + pj + "C = 1.0f - " + pj + "c;");
                const Ex C = Ex::V(pj + "C");
                Ex M[9];
                const float k2[9] = {x * x, x * y, x * z, y * x, y * y, y * z, z * x, z * y, z * z};
                const float k1[9] = {0, -z, y, z, 0, -x, -y, x, 0};
                for (int e = 0; e < 9; e++) {
                    const Ex base = Ex::L(0.f);
                    std::vector<Ex> a = {Ex::L(k2[e]), Ex::L(k1[e])}, b = {C, sn};
                    if (e == 0 || e == 4 || e == 8) { a.push_back(Ex::L(1.f)); b.push_back(cs); }
                    M[e] = g.dot(pj + "m" + std::to_string(e), a, b, &base);
                }
                for (int i = 0; i < 3; i++)
                    for (int c = 0; c < 3; c++)
                        R.m[3 * i + c] = g.dot(pj + "r" + std::to_string(3 * i + c), {Am.m[3 * i], Am.m[3 * i + 1], Am.m[3 * i + 2]},
                                               {M[c], M[3 + c], M[6 + c]});
            }
        } else {   // prismatic: t += R_A * axis * q
            const Ex qv = Ex::V(qn);
            for (int i = 0; i < 3; i++) {
                const Ex d = g.dot(pj + "d" + std::to_string(i), {Am.m[3 * i], Am.m[3 * i + 1], Am.m[3 * i + 2]},
                                   {Ex::L(J.axis[0]), Ex::L(J.axis[1]), Ex::L(J.axis[2])});
                R.m[9 + i] = g.dot(pj + "t" + std::to_string(i), {d}, {qv}, &Am.m[9 + i]);
            }
        }
        X[j] = R;
    }
    // ---- proxy centres of the moving geometries (registers); static ones are literals
    std::vector<Ex> cx(in.ngeoms), cy(in.ngeoms), cz(in.ngeoms);
    for (int gi = 0; gi < in.ngeoms; gi++) {
        const BpGeom &B = in.bpgeoms[gi];
        if (in.geom_parent[gi] == 0) {
            cx[gi] = Ex::L(B.c[0]); cy[gi] = Ex::L(B.c[1]); cz[gi] = Ex::L(B.c[2]);
        } else {
            const Xf &P = X[in.geom_parent[gi]];
            const std::string n = "c" + std::to_string(gi);
            cx[gi] = act_lit(g, n + "x", P, 0, B.c);
            cy[gi] = act_lit(g, n + "y", P, 1, B.c);
            cz[gi] = act_lit(g, n + "z", P, 2, B.c);
        }
    }
    // ---- enclosing capsules (second "certainly free" proxy of elongated shapes): end points, direction
    std::vector<Ex> k0x(in.ngeoms), k0y(in.ngeoms), k0z(in.ngeoms), kux(in.ngeoms), kuy(in.ngeoms), kuz(in.ngeoms);
    std::vector<float> kuu(in.ngeoms, 0.f);
    for (int gi = 0; gi < in.ngeoms; gi++) {
        const BpGeom &B = in.bpgeoms[gi];
        if (!B.has_cap) continue;
        bool used = false;
        for (int p = 0; p < in.npairs && !used; p++)
            used = in.bppairs[p].kind == BP_SPHERES && (in.bppairs[p].a == gi || in.bppairs[p].b == gi);
        if (!used) continue;
        const float du[3] = {B.cap1[0] - B.cap0[0], B.cap1[1] - B.cap0[1], B.cap1[2] - B.cap0[2]};
        kuu[gi] = du[0] * du[0] + du[1] * du[1] + du[2] * du[2];
        const std::string n = "k" + std::to_string(gi);
        if (in.geom_parent[gi] == 0) {
            k0x[gi] = Ex::L(B.cap0[0]); k0y[gi] = Ex::L(B.cap0[1]); k0z[gi] = Ex::L(B.cap0[2]);
            kux[gi] = Ex::L(du[0]); kuy[gi] = Ex::L(du[1]); kuz[gi] = Ex::L(du[2]);
        } else {
            const Xf &P = X[in.geom_parent[gi]];
            k0x[gi] = act_lit(g, n + "x", P, 0, B.cap0); k0y[gi] = act_lit(g, n + "y", P, 1, B.cap0); k0z[gi] = act_lit(g, n + "z", P, 2, B.cap0);
            // direction: rotation only
            const Ex zero = Ex::L(0.f);
            kux[gi] = g.dot(n + "ux", {P.m[0], P.m[1], P.m[2]}, {Ex::L(du[0]), Ex::L(du[1]), Ex::L(du[2])}, &zero);
            kuy[gi] = g.dot(n + "uy", {P.m[3], P.m[4], P.m[5]}, {Ex::L(du[0]), Ex::L(du[1]), Ex::L(du[2])}, &zero);
            kuz[gi] = g.dot(n + "uz", {P.m[6], P.m[7], P.m[8]}, {Ex::L(du[0]), Ex::L(du[1]), Ex::L(du[2])}, &zero);
        }
    }
    // ---- pair tests, group by group (a warp whose lanes are all decided skips the rest)
    auto diff = [&](const std::string &name, const Ex &a, const Ex &b) {
        const Ex one = Ex::L(1.f), neg = Ex::L(-1.f);
        return g.dot(name, {a, b}, {one, neg});
    };
    for (int gi = 0; gi < in.ngroups; gi++) {
        const BpGroup &grp = in.groups[gi];
        if (grp.end <= grp.begin) continue;
        g.line("if (__any_sync(0xffffffffu, alive)) {");
        for (int p = grp.begin; p < grp.end; p++) {
            const BpPair &pr = in.bppairs[p];
            const std::string ps = "p" + std::to_string(p);
            g.line("{");
            Ex d2;
            if (pr.kind == BP_BOX_A || pr.kind == BP_BOX_B) {
                const int gb = pr.kind == BP_BOX_A ? pr.a : pr.b, go = pr.kind == BP_BOX_A ? pr.b : pr.a;
                const BpBox &bx = in.bpboxes[in.bpgeoms[gb].box];
                const Ex dx = diff(ps + "dx", cx[go], Ex::L(bx.t[0])), dy = diff(ps + "dy", cy[go], Ex::L(bx.t[1])),
                         dz = diff(ps + "dz", cz[go], Ex::L(bx.t[2]));
                Ex e[3];
                for (int k = 0; k < 3; k++) {
                    const Ex l = g.dot(ps + "l" + std::to_string(k), {dx, dy, dz}, {Ex::L(bx.R[k]), Ex::L(bx.R[3 + k]), Ex::L(bx.R[6 + k])});
                    if (l.lit) e[k] = Ex::L(fmaxf(fabsf(l.v) - bx.h[k], 0.f));
                    else {
                        g.line("const float " + ps + "e" + std::to_string(k) + " = fmaxf(fabsf(" + l.str() + ") - " + Ex::L(bx.h[k]).str() + ", 0.f);");
                        e[k] = Ex::V(ps + "e" + std::to_string(k));
                    }
                }
                d2 = g.dot(ps + "d2", {e[0], e[1], e[2]}, {e[0], e[1], e[2]});
            } else {
                const Ex dx = diff(ps + "dx", cx[pr.a], cx[pr.b]), dy = diff(ps + "dy", cy[pr.a], cy[pr.b]), dz = diff(ps + "dz", cz[pr.a], cz[pr.b]);
                d2 = g.dot(ps + "d2", {dx, dy, dz}, {dx, dy, dz});
            }
            const std::string P = std::to_string(p);
            g.line("const float d2v = " + d2.str() + ";");
            g.line("const bool sure = alive && d2v <= A.thr[" + P + "][1];");
            std::string cap_free;
            if (pr.kind == BP_SPHERES) {
                const bool ca = in.bpgeoms[pr.a].has_cap && kuu[pr.a] > 0.f, cb2 = in.bpgeoms[pr.b].has_cap && kuu[pr.b] > 0.f;
                auto f = [](float x) { return Ex::L(x).str(); };
                if (ca && cb2) {
                    g.line("const float dk = seg_seg2(" + k0x[pr.a].str() + ", " + k0y[pr.a].str() + ", " + k0z[pr.a].str() + ", " + kux[pr.a].str() + ", " +
                           kuy[pr.a].str() + ", " + kuz[pr.a].str() + ", " + f(kuu[pr.a]) + ", " + k0x[pr.b].str() + ", " + k0y[pr.b].str() + ", " +
                           k0z[pr.b].str() + ", " + kux[pr.b].str() + ", " + kuy[pr.b].str() + ", " + kuz[pr.b].str() + ", " + f(kuu[pr.b]) + ");");
                    cap_free = " && !(dk > A.thr[" + P + "][2])";
                } else if (ca || cb2) {
                    const int gc = ca ? pr.a : pr.b, go = ca ? pr.b : pr.a;
                    g.line("const float dk = pt_seg2(" + cx[go].str() + ", " + cy[go].str() + ", " + cz[go].str() + ", " + k0x[gc].str() + ", " +
                           k0y[gc].str() + ", " + k0z[gc].str() + ", " + kux[gc].str() + ", " + kuy[gc].str() + ", " + kuz[gc].str() + ", " +
                           f(1.0f / kuu[gc]) + ");");
                    cap_free = " && !(dk > A.thr[" + P + "][2])";
                }
            }
            g.line("const bool und = alive && !sure && !(d2v > A.thr[" + P + "][0])" + cap_free + ";");
            g.line("if (sure) { state_hit = true; alive = false; }");
            g.line("w" + std::to_string(p >> 5) + " |= und ? " + std::to_string(1u << (p & 31)) + "u : 0u;");
            g.line("}");
        }
        g.line("}");
    }
    s += g.out;
    s += "    }\n";
    // ---- epilogue: append the undecided items to their pairs' bins.  Every thread holds its items
    //      as bits of w0..; two passes over its OWN bits with shared-memory atomics: count per pair,
    //      one global reservation per (block, pair), then rank + store.  (The generic kernel's
    //      warp-ballot version walks every pair present in the warp twice, ~800 instructions per
    //      warp; here a thread pays only for its ~1.3 items.)
    s += "    if (state_hit) {";
    for (int w = 0; w < NW; w++) s += " w" + std::to_string(w) + " = 0u;";
    s += " }\n";
    for (int w = 0; w < NW; w++)
        s += "    { uint32_t b = w" + std::to_string(w) + "; while (b) { const int k = __ffs(b) - 1; b &= b - 1; atomicAdd(&scnt[" +
             std::to_string(32 * w) + " + k], 1u); } }\n";
    s += "    __syncthreads();\n"
         "    for (int p = tid; p < NP; p += BT) {\n"
         "        const uint32_t c = scnt[p];\n"
         "        sbase[p] = c ? atomicAdd(A.bin_count + p, c) : 0u;\n"
         "        scnt[p] = 0u;\n"
         "    }\n"
         "    __syncthreads();\n";
    for (int w = 0; w < NW; w++)
        s += "    { uint32_t b = w" + std::to_string(w) + "; while (b) { const int k = __ffs(b) - 1; b &= b - 1; const int p = " +
             std::to_string(32 * w) + " + k; const uint32_t r = atomicAdd(&scnt[p], 1u); A.bin_state[(int64_t)p * A.cap + sbase[p] + r] = (int32_t)state; } }\n";
    s += "    if (valid && state_hit) A.hit[group] = 1;\n}\n";
    return s;
}

// ---- NVRTC: source -> cubin for sm_XYa, cached per process by source text
// NVRTC is loaded at run time (dlopen): the engine library has no link-time dependency on it, and a
// machine without it simply keeps the generic kernel.
struct NvrtcApi {
    nvrtcResult (*createProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*compileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*getProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*getCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*destroyProgram)(nvrtcProgram *) = nullptr;
    bool ok = false;
};
inline const NvrtcApi &nvrtc_api() {
    static NvrtcApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = nullptr;
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "libnvrtc.so.13"})
            if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
        if (h) {
            auto get = [&](const char *sym, void **fn) { *fn = dlsym(h, sym); return *fn != nullptr; };
            api.ok = get("nvrtcCreateProgram", (void **)&api.createProgram) && get("nvrtcCompileProgram", (void **)&api.compileProgram) &&
                     get("nvrtcGetProgramLogSize", (void **)&api.getProgramLogSize) && get("nvrtcGetProgramLog", (void **)&api.getProgramLog) &&
                     get("nvrtcGetCUBINSize", (void **)&api.getCUBINSize) && get("nvrtcGetCUBIN", (void **)&api.getCUBIN) &&
                     get("nvrtcDestroyProgram", (void **)&api.destroyProgram);
        }
    }
    return api;
}

inline int compile(const std::string &src, int cc_major, int cc_minor, std::vector<char> &cubin, std::string &log) {
    static std::mutex mu;
    static std::map<std::string, std::vector<char>> cache;
    std::lock_guard<std::mutex> lock(mu);
    const NvrtcApi &rt = nvrtc_api();
    if (!rt.ok) { log = "NVRTC (libnvrtc.so) is not available"; return -1; }
    char arch[64];
    snprintf(arch, sizeof arch, "--gpu-architecture=sm_%d%d%s", cc_major, cc_minor, cc_major >= 9 ? "a" : "");
    const std::string key = std::string(arch) + "\n" + src;
    auto it = cache.find(key);
    if (it != cache.end()) { cubin = it->second; return 0; }
    nvrtcProgram prog;
    if (rt.createProgram(&prog, src.c_str(), "pbp_bp_spec.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return -1; }
    const char *opts[] = {arch, "--fmad=false", "-lineinfo", "--std=c++17"};
    const nvrtcResult rc = rt.compileProgram(prog, 4, opts);
    size_t ls = 0;
    rt.getProgramLogSize(prog, &ls);
    if (ls > 1) { log.resize(ls); rt.getProgramLog(prog, &log[0]); }
    if (rc != NVRTC_SUCCESS) { rt.destroyProgram(&prog); return -1; }
    size_t cs = 0;
    if (rt.getCUBINSize(prog, &cs) != NVRTC_SUCCESS || cs == 0) { rt.destroyProgram(&prog); log += " (no cubin)"; return -1; }
    cubin.resize(cs);
    rt.getCUBIN(prog, cubin.data());
    rt.destroyProgram(&prog);
    cache[key] = cubin;
    return 0;
}

}  // namespace jit
}  // namespace pbp
