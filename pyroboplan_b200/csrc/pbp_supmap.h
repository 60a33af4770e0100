// Exact support maps for convex hulls (host-side builder, plain C++).
//
// The support function of a hull, argmax_v <v, d>, only depends on the direction of d.  The
// direction sphere is cut into 6 * K * K cells (cube faces, K x K squares in the (u, w) = d_b/|d_a|,
// d_c/|d_a| chart of the face with major axis a).  For every cell the builder lists EVERY vertex
// that maximises <v, d> for some d in the (slightly inflated) cell -- a vertex is kept iff the
// region of the cell where it beats all other vertices is non-empty, decided by clipping the cell
// rectangle with the half-planes f_v >= f_x (f linear in (u, w)).  Scanning a cell's list is
// therefore exactly equivalent to scanning the whole hull, at ~3 instead of ~150 vertices.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pbp {

// 16 cells per face edge: ~2 candidates per cell (8: ~3, with long tails that made the lanes of a
// warp wait for the longest list; measured narrowphase 170 -> 160 us).  20 / 24 gain nothing more.
#ifndef PBP_SUPMAP_K
#define PBP_SUPMAP_K 16
#endif
constexpr int SUPMAP_K = PBP_SUPMAP_K;                   // cells per cube-face edge
constexpr int SUPMAP_CELLS = 6 * SUPMAP_K * SUPMAP_K;    // 1536

struct SupMap {
    std::vector<uint16_t> cell_off;   // [SUPMAP_CELLS + 1] offsets into list
    std::vector<uint16_t> list;       // candidate vertex indices
};

namespace supmap_detail {
struct P2 { double u, w; };

// clip polygon with a*u + b*w + c >= 0
inline void clip(std::vector<P2> &poly, double a, double b, double c, std::vector<P2> &tmp) {
    tmp.clear();
    const size_t n = poly.size();
    for (size_t i = 0; i < n; i++) {
        const P2 &p = poly[i], &q = poly[(i + 1) % n];
        const double fp = a * p.u + b * p.w + c, fq = a * q.u + b * q.w + c;
        if (fp >= 0.0) tmp.push_back(p);
        if ((fp >= 0.0) != (fq >= 0.0)) {
            const double t = fp / (fp - fq);
            tmp.push_back(P2{p.u + t * (q.u - p.u), p.w + t * (q.w - p.w)});
        }
    }
    poly.swap(tmp);
}
}  // namespace supmap_detail

// verts: n x 3 (x, y, z), float precision values as the device sees them.
inline SupMap build_support_map(const float *verts, int n) {
    using supmap_detail::P2;
    SupMap m;
    m.cell_off.assign(SUPMAP_CELLS + 1, 0);
    const int K = SUPMAP_K;
    const double inflate = 1e-4;    // cells overlap a little: FP32 rounding of (u, w) on device cannot miss a maximiser
    double scale = 0.0;
    for (int i = 0; i < 3 * n; i++) scale = std::max(scale, (double)std::fabs(verts[i]));
    const double eps = 1e-9 * std::max(scale, 1e-30);   // inclusive comparisons: ties keep both vertices
    std::vector<P2> poly, tmp;
    std::vector<double> centre_val(n);
    std::vector<int> order(n);
    for (int face = 0; face < 6; face++) {
        const int a = face >> 1, b = (a + 1) % 3, c = (a + 2) % 3;
        const double s = (face & 1) ? -1.0 : 1.0;
        for (int iu = 0; iu < K; iu++)
            for (int iw = 0; iw < K; iw++) {
                const int cell = face * K * K + iu * K + iw;
                const double u0 = -1.0 + 2.0 * iu / K - inflate, u1 = -1.0 + 2.0 * (iu + 1) / K + inflate;
                const double w0 = -1.0 + 2.0 * iw / K - inflate, w1 = -1.0 + 2.0 * (iw + 1) / K + inflate;
                const double uc = 0.5 * (u0 + u1), wc = 0.5 * (w0 + w1);
                // f_x(u, w) = s * x_a + u * x_b + w * x_c
                for (int x = 0; x < n; x++) {
                    centre_val[x] = s * verts[3 * x + a] + uc * verts[3 * x + b] + wc * verts[3 * x + c];
                    order[x] = x;
                }
                std::sort(order.begin(), order.end(), [&](int p, int q) { return centre_val[p] > centre_val[q]; });
                m.cell_off[cell] = (uint16_t)m.list.size();
                for (int v = 0; v < n; v++) {
                    poly.assign({P2{u0, w0}, P2{u1, w0}, P2{u1, w1}, P2{u0, w1}});
                    const double va = verts[3 * v + a], vb = verts[3 * v + b], vc = verts[3 * v + c];
                    for (int k = 0; k < n && !poly.empty(); k++) {
                        const int x = order[k];
                        if (x == v) continue;
                        // f_v - f_x >= -eps
                        supmap_detail::clip(poly, vb - verts[3 * x + b], vc - verts[3 * x + c], s * (va - verts[3 * x + a]) + eps, tmp);
                    }
                    if (!poly.empty()) m.list.push_back((uint16_t)v);
                }
            }
    }
    m.cell_off[SUPMAP_CELLS] = (uint16_t)m.list.size();
    for (int i = 0; i < 4; i++) m.list.push_back(m.list.empty() ? (uint16_t)0 : m.list.back());  // the device reads 4 entries per trip
    return m;
}

}  // namespace pbp
