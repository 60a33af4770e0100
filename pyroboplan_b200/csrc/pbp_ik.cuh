// Batched damped-least-squares differential IK (one "try" per launch, thread per target).
//
// Restates one try of DifferentialIk.solve (/root/reference/src/pyroboplan/ik/differential_ik.py:208-298):
//   FK of the target frame -> error = -log6(target^-1 * current) -> converged? (wrap to (-pi, pi],
//   joint limits) : LOCAL frame Jacobian J -> (J W J^T + damping^2 I) y = error ->
//   q += alpha * W J^T y,  alpha = min_step + (1 - |e| / |e0|) (max_step - min_step).
// The 6 x nv system is tiny and the damping makes it ill-conditioned near singularities, so this
// kernel (unlike the collision kernels) works in FP64 END TO END: states, targets and the model tables
// of the chain are doubles (a 1e-8 perturbation of a joint placement is amplified by up to
// 1/damping^2 = 1e6 when an iterate crosses a singularity); B200 runs FP64 at half the FP32 rate.
//
// FK and Jacobian come from ONE backward sweep over the joints on the frame's chain, no world
// transforms stored: with X = pose of the frame in joint k's frame, the LOCAL Jacobian column of a
// revolute joint with axis a is [R_X^T (a x t_X); R_X^T a] (prismatic: [R_X^T a; 0]), and
// X <- (P_k * motion_k(q_k)) * X moves one joint towards the root; after the root, X = oMf.
#pragma once
#include "pbp_model.cuh"

namespace pbp {

constexpr int IK_MAX_PATH = 32;

struct IkOptions {
    int max_iters;
    double tol_translation, tol_rotation, damping, min_step, max_step;
};

// One joint of the chain root -> frame (float64 copy of the model tables)
struct IkJoint {
    double P[12];     // placement in the parent joint frame
    double axis[3];
    double w;         // diagonal of W for this coordinate, 0 when the coordinate is not active
    int32_t type, idx_q;
};

struct IkTables {
    IkJoint joint[IK_MAX_PATH];
    double frame[12];                         // placement of the frame in its parent joint
    double q_lower[PBP_MAX_NQ], q_upper[PBP_MAX_NQ];
    unsigned long long active_mask;           // coordinates that may move
    unsigned long long path_mask;             // coordinates of the joints on the chain
    int32_t npath, nq;
};

__device__ __forceinline__ void d_se3_mul(const double *A, const double *B, double *C) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double a0 = A[3 * i], a1 = A[3 * i + 1], a2 = A[3 * i + 2];
        C[3 * i + 0] = a0 * B[0] + a1 * B[3] + a2 * B[6];
        C[3 * i + 1] = a0 * B[1] + a1 * B[4] + a2 * B[7];
        C[3 * i + 2] = a0 * B[2] + a1 * B[5] + a2 * B[8];
        C[9 + i] = a0 * B[9] + a1 * B[10] + a2 * B[11] + A[9 + i];
    }
}

// -log6(E): minus the se(3) logarithm of E = (R, p) as [linear; angular] (Pinocchio's Motion order)
__device__ __forceinline__ void d_neg_log6(const double *E, double *e) {
    const double vx = E[7] - E[5], vy = E[2] - E[6], vz = E[3] - E[1];  // vee(R - R^T)
    const double s = 0.5 * sqrt(vx * vx + vy * vy + vz * vz);           // sin(theta)
    const double c = 0.5 * (E[0] + E[4] + E[8] - 1.0);                  // cos(theta)
    const double theta = atan2(s, c);
    double w[3];
    if (c > -0.99) {
        const double k = theta > 1e-8 ? theta / (2.0 * s) : 0.5 * (1.0 + theta * theta / 6.0);
        w[0] = k * vx; w[1] = k * vy; w[2] = k * vz;
    } else {
        // close to pi: axis from the symmetric part, sign from the antisymmetric part
        const double d0 = E[0], d1 = E[4], d2 = E[8];
        const int kx = (d0 >= d1 && d0 >= d2) ? 0 : (d1 >= d2 ? 1 : 2);
        const double dk = kx == 0 ? d0 : (kx == 1 ? d1 : d2);
        double ax[3];
        const double xk = sqrt(fmax((dk - c) / (1.0 - c), 0.0));
        const double inv = 1.0 / (2.0 * xk * (1.0 - c));
        ax[kx] = xk;
        ax[(kx + 1) % 3] = (E[3 * ((kx + 1) % 3) + kx] + E[3 * kx + (kx + 1) % 3]) * inv;
        ax[(kx + 2) % 3] = (E[3 * ((kx + 2) % 3) + kx] + E[3 * kx + (kx + 2) % 3]) * inv;
        const double sg = (ax[0] * vx + ax[1] * vy + ax[2] * vz) < 0.0 ? -1.0 : 1.0;
        const double nrm = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        for (int i = 0; i < 3; i++) w[i] = sg * theta * ax[i] / nrm;
    }
    const double t2 = theta * theta;
    double alpha, beta;
    if (theta < 1e-4) {
        alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
        beta = 1.0 / 12.0 + t2 / 720.0;
    } else {
        const double st = sin(theta), ct = cos(theta);
        alpha = theta * st / (2.0 * (1.0 - ct));
        beta = 1.0 / t2 - st / (2.0 * theta * (1.0 - ct));
    }
    const double px = E[9], py = E[10], pz = E[11];
    const double wp = w[0] * px + w[1] * py + w[2] * pz;
    const double cx = w[1] * pz - w[2] * py, cy = w[2] * px - w[0] * pz, cz = w[0] * py - w[1] * px;
    e[0] = -(alpha * px - 0.5 * cx + beta * wp * w[0]);
    e[1] = -(alpha * py - 0.5 * cy + beta * wp * w[1]);
    e[2] = -(alpha * pz - 0.5 * cz + beta * wp * w[2]);
    e[3] = -w[0]; e[4] = -w[1]; e[5] = -w[2];
}


// status bits: 1 = pose error within tolerance, 2 = (after wrapping) within the joint limits
template <int MAXP>
__global__ void __launch_bounds__(128)
k_ik_iterate(const IkTables *__restrict__ tables, const double *__restrict__ targets, double *__restrict__ q_io, int64_t n,
             IkOptions opt, const double *__restrict__ nullspace, double *__restrict__ e0_io, int32_t *__restrict__ status,
             int32_t *__restrict__ iters_out, unsigned long long *__restrict__ work_cursor) {
    __shared__ IkTables T;
    {
        const int words = (int)(sizeof(IkTables) / 8);
        const double *src = reinterpret_cast<const double *>(tables);
        double *dst = reinterpret_cast<double *>(&T);
        for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    // Persistent lanes: a lane runs ONE iteration per trip for its current target and takes the next
    // target from a global cursor when that one is finished.  Targets need anything from 0 to
    // max_iters iterations (mean ~25, and ~1 % never converge): with a fixed thread <-> target mapping
    // every warp would run as long as its slowest lane.
    const int npath = T.npath, nq = T.nq;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int64_t i = -1;
    bool have = false, more = true;
    double *q = q_io;
    double qp[MAXP], J[MAXP][6];
    double Tt[12];
    double e0 = 0.0;
    int st = 0, it = 0;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, !have);
        if (need && more) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work_cursor, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + __popc(need) >= (unsigned long long)n) more = false;
            if (!have) {
                i = (int64_t)base + __popc(need & lt_mask);
                if (i < n) {
                    q = q_io + i * nq;
#pragma unroll
                    for (int k = 0; k < MAXP; k++) qp[k] = k < npath ? q[T.joint[k].idx_q] : 0.0;
#pragma unroll
                    for (int k = 0; k < 12; k++) Tt[k] = targets[i * 12 + k];
                    e0 = e0_io[i];
                    st = 0;
                    it = 0;
                    have = true;
                }
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (!have) continue;
        bool done = it >= opt.max_iters;
        if (!done) {
        double X[12];
#pragma unroll
        for (int k = 0; k < 12; k++) X[k] = T.frame[k];
#pragma unroll
        for (int k = MAXP - 1; k >= 0; k--) {
            if (k >= npath) continue;
            const IkJoint &Jt = T.joint[k];
            const double ax = Jt.axis[0], ay = Jt.axis[1], az = Jt.axis[2];
            // LOCAL Jacobian column through X = (R, t): the frame seen from joint k
            double lx, ly, lz, rx = 0.0, ry = 0.0, rz = 0.0;
            double Mo[12];
#pragma unroll
            for (int m = 0; m < 12; m++) Mo[m] = (m == 0 || m == 4 || m == 8) ? 1.0 : 0.0;
            if (Jt.type == PBP_JOINT_REVOLUTE) {
                const double cx = ay * X[11] - az * X[10], cy = az * X[9] - ax * X[11], cz = ax * X[10] - ay * X[9];  // a x t
                lx = X[0] * cx + X[3] * cy + X[6] * cz; ly = X[1] * cx + X[4] * cy + X[7] * cz; lz = X[2] * cx + X[5] * cy + X[8] * cz;
                rx = X[0] * ax + X[3] * ay + X[6] * az; ry = X[1] * ax + X[4] * ay + X[7] * az; rz = X[2] * ax + X[5] * ay + X[8] * az;
                double s, c;
                sincos(qp[k], &s, &c);
                const double C = 1.0 - c;
                Mo[0] = c + ax * ax * C;      Mo[1] = ax * ay * C - az * s; Mo[2] = ax * az * C + ay * s;
                Mo[3] = ay * ax * C + az * s; Mo[4] = c + ay * ay * C;      Mo[5] = ay * az * C - ax * s;
                Mo[6] = az * ax * C - ay * s; Mo[7] = az * ay * C + ax * s; Mo[8] = c + az * az * C;
            } else {
                lx = X[0] * ax + X[3] * ay + X[6] * az; ly = X[1] * ax + X[4] * ay + X[7] * az; lz = X[2] * ax + X[5] * ay + X[8] * az;
                Mo[9] = ax * qp[k]; Mo[10] = ay * qp[k]; Mo[11] = az * qp[k];
            }
            J[k][0] = lx; J[k][1] = ly; J[k][2] = lz; J[k][3] = rx; J[k][4] = ry; J[k][5] = rz;
            // X <- (P_k * motion_k(q_k)) * X
            double L[12], Y[12];
            d_se3_mul(Jt.P, Mo, L);
            d_se3_mul(L, X, Y);
#pragma unroll
            for (int m = 0; m < 12; m++) X[m] = Y[m];
        }
        // E = target^-1 * current
        double E[12];
        {
            const double dx = X[9] - Tt[9], dy = X[10] - Tt[10], dz = X[11] - Tt[11];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const double a0 = Tt[r], a1 = Tt[3 + r], a2 = Tt[6 + r];
                E[3 * r + 0] = a0 * X[0] + a1 * X[3] + a2 * X[6];
                E[3 * r + 1] = a0 * X[1] + a1 * X[4] + a2 * X[7];
                E[3 * r + 2] = a0 * X[2] + a1 * X[5] + a2 * X[8];
                E[9 + r] = a0 * dx + a1 * dy + a2 * dz;
            }
        }
        double e[6];
        d_neg_log6(E, e);
        const double et = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]), er = sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
        if (et < opt.tol_translation && er < opt.tol_rotation) {
            st = 1;
            done = true;
        }
        if (!done) {
        // null-space term (reference :274-283): the right-hand side becomes e - J ns
        if (nullspace) {
#pragma unroll
            for (int k = 0; k < MAXP; k++) {
                if (k >= npath || !(T.joint[k].w > 0.0)) continue;
                const double nsk = nullspace[i * nq + T.joint[k].idx_q];
#pragma unroll
                for (int r = 0; r < 6; r++) e[r] -= J[k][r] * nsk;
            }
        }
        // A = J W J^T + damping^2 I (lower triangle), Cholesky, solve A y = e
        double A[6][6];
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c2 = 0; c2 <= r; c2++) {
                double sacc = r == c2 ? opt.damping * opt.damping : 0.0;
#pragma unroll
                for (int k = 0; k < MAXP; k++)
                    if (k < npath) sacc += T.joint[k].w * J[k][r] * J[k][c2];
                A[r][c2] = sacc;
            }
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
            for (int c2 = 0; c2 <= r; c2++) {
                double sacc = A[r][c2];
#pragma unroll
                for (int m = 0; m < c2; m++) sacc -= A[r][m] * A[c2][m];
                A[r][c2] = r == c2 ? sqrt(fmax(sacc, 1e-300)) : sacc / A[c2][c2];
            }
        }
        double y[6];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double sacc = e[r];
#pragma unroll
            for (int m = 0; m < r; m++) sacc -= A[r][m] * y[m];
            y[r] = sacc / A[r][r];
        }
#pragma unroll
        for (int r = 5; r >= 0; r--) {
            double sacc = y[r];
#pragma unroll
            for (int m = r + 1; m < 6; m++) sacc -= A[m][r] * y[m];
            y[r] = sacc / A[r][r];
        }
        const double en = sqrt(et * et + er * er);   // norm of the pose error itself (before the null-space shift)
        if (!(e0 > 0.0)) e0 = en;
        const double alpha = opt.min_step + (1.0 - en / e0) * (opt.max_step - opt.min_step);
        bool bad = false;
#pragma unroll
        for (int k = 0; k < MAXP; k++) {
            if (k >= npath) continue;
            double dq = 0.0;
#pragma unroll
            for (int r = 0; r < 6; r++) dq += J[k][r] * y[r];
            double step = T.joint[k].w * dq;
            if (nullspace && T.joint[k].w > 0.0) step += nullspace[i * nq + T.joint[k].idx_q];
            qp[k] += alpha * step;
            if (isinf(qp[k])) bad = true;
        }
        if (nullspace) {   // active coordinates off the chain have zero Jacobian columns: only the null-space term moves them
            for (int k = 0; k < nq; k++)
                if (((T.active_mask >> k) & 1ull) && !((T.path_mask >> k) & 1ull)) {
                    q[k] += alpha * nullspace[i * nq + k];
                    if (isinf(q[k])) bad = true;
                }
        }
        it++;
        if (bad) done = true;
        }   // step
        }   // iteration
        if (!done) continue;
        // ---- this target is finished: write back, free the lane
#pragma unroll
        for (int k = 0; k < MAXP; k++)
            if (k < npath) q[T.joint[k].idx_q] = qp[k];
        if (st & 1) {
            // wrap EVERY coordinate to [-pi, pi) as the reference does before its limit check (:221-222)
            const double PI = 3.14159265358979323846;
            bool within = true;
            for (int k = 0; k < nq; k++) {
                double v = q[k] + PI;
                v = v - (2.0 * PI) * floor(v / (2.0 * PI)) - PI;   // Python's % for a positive modulus
                q[k] = v;
                if (v < T.q_lower[k] || v > T.q_upper[k]) within = false;
            }
            if (within) st |= 2;
        }
        e0_io[i] = e0;
        status[i] = st;
        if (iters_out) iters_out[i] = it;
        have = false;
    }
}

}  // namespace pbp
