// pam_core.h -- per-item numerics of the part-aware measurement path (host/device inline).
//
// Every function here is the arithmetic ONE thread performs for ONE work item (a joint, a view
// pair, a track/detection pair, one assignment problem).  The file is plain C++ so that the
// very same source is compiled by nvcc into the sm_100a kernels (pam_kernels.cu) and by g++
// into the host-side debugging harness under tests/hostemu/ (never shipped, never loaded by
// the package).
//
// Reference semantics followed (paths under /root/reference/src):
//   np_sum / np_mean            numpy add.reduce order (pairwise_sum in numpy/_core/src/umath/loops_utils.h.src)
//   epi_line_* / epi_dist_*     utils/matching.py:115-151 (float64 form), :50-91 + cv::computeCorrespondEpilines
//   ray_point_distance          utils/matching.py:10-17 + utils/calculate.py:26-32
//   greedy_update / greedy_init utils/matching.py:243-295
//   dlt_*                       utils/construction.py:89-114  (rows, weights; the SVD is replaced by
//                               a streaming Givens QR + one-sided Jacobi SVD of the 4x4 factor)
//   lsap_solve                  scipy.optimize.linear_sum_assignment (rectangular_lsap.cpp, Crouse 2016)
//   reflect_index / gauss_last  scipy.ndimage.gaussian_filter1d(mode='reflect') last output sample,
//                               tracking/IterativeTracker.py:371-383
#pragma once
#include <math.h>
#include <string.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PAM_HD __host__ __device__ __forceinline__
#define PAM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PAM_HD inline
#define PAM_HD_NOINLINE
#endif

// The tracker kernel is a long straight-line program executed once per frame; keeping its hot part
// inside the 32 KB L1.5 instruction cache matters more than saving loop overhead, so loops with
// run-time trip counts are kept rolled.
#if defined(__CUDACC__)
#define PAM_NOUNROLL _Pragma("unroll 1")
#define PAM_UNROLL2 _Pragma("unroll 2")
#define PAM_UNROLL4 _Pragma("unroll 4")
#define PAM_PRAGMA_(x) _Pragma(#x)
#define PAM_UNROLL_N(n) PAM_PRAGMA_(unroll (n))
#else
#define PAM_NOUNROLL
#define PAM_UNROLL2
#define PAM_UNROLL4
#define PAM_UNROLL_N(n)
#endif

#define PAM_MAX_V 8        // cameras per rig handled by the stateful tracker
#define PAM_MAX_TRK 32     // track slots per sequence (run-time value: pam_config.max_tracks)
#define PAM_MAX_D 16       // detections per camera per frame
#define PAM_MAX_J 32       // joints
#define PAM_RECENT 5       // newest history entries the smoothing / velocity step keeps in registers
#define PAM_MAX_HYP 64     // person hypotheses during new-track initialisation (run-time value: min(V * D, 64))
#define PAM_LSAP_N 64      // largest assignment problem solved inside the tracker (tracks, hypotheses <= 64)
#define PAM_HIST 12        // smoothed-pose history ring (max_age + 2 <= PAM_HIST)
#define PAM_MAX_RADIUS 8   // Gaussian radius int(4 sigma + 0.5)
#define PAM_MAX_AGEW 8     // stale-view window + 1
#define PAM_VLIST (2 + PAM_MAX_V)   // bytes per reported track of the optional view-list output

namespace pam {

// x / d for 0 <= x < 2^20 and small d, with inv = 1.0f / d: exact, 4 instructions instead of ~20
PAM_HD int fast_div(int x, float inv) { return (int)(((float)x + 0.5f) * inv); }

PAM_HD int popcount32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// index of the lowest set bit (x != 0)
PAM_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// ------------------------------------------------------------------------------------------
// numpy reduction order
// ------------------------------------------------------------------------------------------
// n <= 128: the non-recursive leaf of numpy's pairwise_sum (all on-device reductions: J <= 32, V <= 31)
template <class T>
PAM_HD T np_sum(const T* a, int n, int stride = 1) {
    if (n < 8) {
        T r = (T)0;
        PAM_NOUNROLL for (int i = 0; i < n; ++i) r += a[i * stride];
        return r;
    }
    T r0 = a[0], r1 = a[stride], r2 = a[2 * stride], r3 = a[3 * stride];
    T r4 = a[4 * stride], r5 = a[5 * stride], r6 = a[6 * stride], r7 = a[7 * stride];
    int i = 8;
    PAM_NOUNROLL for (; i < n - (n % 8); i += 8) {
        r0 += a[(i + 0) * stride]; r1 += a[(i + 1) * stride];
        r2 += a[(i + 2) * stride]; r3 += a[(i + 3) * stride];
        r4 += a[(i + 4) * stride]; r5 += a[(i + 5) * stride];
        r6 += a[(i + 6) * stride]; r7 += a[(i + 7) * stride];
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    PAM_NOUNROLL for (; i < n; ++i) res += a[i * stride];
    return res;
}

// running 8-lane accumulator reproducing np_sum for a stream of n <= 128 values whose count is
// known up front (used where the addends are produced on the fly and never stored)
template <class T>
struct NpSumStream {
    T r[8];
    T res;
    int n, i, nblk;
    PAM_HD void begin(int count) {
        n = count; i = 0; nblk = count - (count % 8); res = (T)0;
        for (int k = 0; k < 8; ++k) r[k] = (T)0;
    }
    PAM_HD void push(T x) {
        if (n < 8) { res += x; }
        else if (i < nblk) {
            int k = i & 7;
            if (i < 8) r[k] = x; else r[k] += x;
            if (i == nblk - 1) res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        } else { res += x; }
        ++i;
    }
    PAM_HD T total() const { return res; }
};

// get_believe (utils/calculate.py:8-14): mean of the confidences that are >= 0, summed in numpy's
// order.  q = (v, u, conf) triples of one detection.  Common case (no negative confidence): one strided
// pass; otherwise the selected values are streamed through the same accumulator pattern.
PAM_HD double mean_confidence(const float* q, int J) {
    int nk = 0;
    PAM_NOUNROLL for (int j = 0; j < J; ++j) nk += (q[j * 3 + 2] >= 0.0f) ? 1 : 0;
    if (nk == J) {
        double res;
        if (J < 8) {
            res = 0.0;
            PAM_NOUNROLL for (int j = 0; j < J; ++j) res += (double)q[j * 3 + 2];
        } else {
            double r0 = q[2], r1 = q[5], r2 = q[8], r3 = q[11], r4 = q[14], r5 = q[17], r6 = q[20], r7 = q[23];
            int i = 8;
            PAM_NOUNROLL for (; i < J - (J % 8); i += 8) {
                const float* p = q + i * 3 + 2;
                r0 += (double)p[0]; r1 += (double)p[3]; r2 += (double)p[6]; r3 += (double)p[9];
                r4 += (double)p[12]; r5 += (double)p[15]; r6 += (double)p[18]; r7 += (double)p[21];
            }
            res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
            PAM_NOUNROLL for (; i < J; ++i) res += (double)q[i * 3 + 2];
        }
        return res / (double)J;
    }
    NpSumStream<double> acc;
    acc.begin(nk);
    PAM_NOUNROLL for (int j = 0; j < J; ++j)
        if (q[j * 3 + 2] >= 0.0f) acc.push((double)q[j * 3 + 2]);
    return acc.total() / (double)nk;              // 0/0 -> NaN like np.mean([])
}

// ------------------------------------------------------------------------------------------
// epipolar geometry
// ------------------------------------------------------------------------------------------
// F is the 3x3 (row-major) fundamental matrix cams[a].F[cid_b]:  x_a^T F x_b = 0,  x = (u, v, 1).

// Short FP64 special functions for the device: one MUFU seed (rsqrt.approx / rcp.approx, ~2^-23..2^-26
// relative) + two Newton steps = full double precision to 1-2 ulp in 5-9 instructions, versus the
// 15-30 instructions (plus slow-path branches) of the IEEE-rounded library routines.  They keep the
// hot loop small enough for the instruction cache.  Arguments are positive normal numbers wherever
// these are used (squared pixel norms, homogeneous depths, pivots); the plain C versions serve the
// host harness.
PAM_HD double rsqrt_f64(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
#else
    return 1.0 / sqrt(x);
#endif
}
PAM_HD double rcp_f64(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}
// sqrt for x >= 0 (0 -> 0)
PAM_HD double sqrt_f64(double x) {
#if defined(__CUDA_ARCH__)
    if (!(x > 0.0)) return 0.0;
    const double y = rsqrt_f64(x);
    double s = x * y;
    const double r = fma(-s, s, x);
    return fma(r, 0.5 * y, s);
#else
    return sqrt(x);
#endif
}

// float64 form of utils/matching.py:136-146: line l = F^T x_a in image b, distance of x_b to it.
// The reference normalises l by ||l[:2]|| (0 -> 1) and then divides by sqrt(l0^2 + l1^2) = 1 again;
// here both collapse into one reciprocal square root (differs by rounding only).
PAM_HD double epi_dist_f64(const double* F, double ua, double va, double ub, double vb) {
    double l0 = F[0] * ua + F[3] * va + F[6];
    double l1 = F[1] * ua + F[4] * va + F[7];
    double l2 = F[2] * ua + F[5] * va + F[8];
    double n2 = l0 * l0 + l1 * l1;
    double inv = (n2 == 0.0) ? 1.0 : rsqrt_f64(n2);
    return fabs(ub * l0 + vb * l1 + l2) * inv;
}

// Raw form of the same: numerator t = l . x_b and squared norm n2 = l0^2 + l1^2 of the line F^T x_a, so
// that "distance < thr" can be decided as t^2 < thr^2 n2 without a square root.
PAM_HD void epi_raw_f64(const double* F, double ua, double va, double ub, double vb, double& t, double& n2) {
    const double l0 = F[0] * ua + F[3] * va + F[6];
    const double l1 = F[1] * ua + F[4] * va + F[7];
    const double l2 = F[2] * ua + F[5] * va + F[8];
    n2 = l0 * l0 + l1 * l1;
    t = ub * l0 + vb * l1 + l2;
}

// cv::computeCorrespondEpilines arithmetic (double path): (a,b,c) = M x, nu = a^2+b^2,
// nu = nu ? 1/sqrt(nu) : 1, scaled;  then utils/matching.py:82-83: |x . l| / sqrt(l0^2 + l1^2)
// (= 1 after the scaling; for nu == 0 the reference divides by 0 and so does this).
// transposed = false: M = F ;  true: M = F^T.
PAM_HD double epi_dist_cv(const double* F, bool transposed, double xs, double ys, double xt, double yt) {
    double a, b, c;
    if (!transposed) {
        a = F[0] * xs + F[1] * ys + F[2];
        b = F[3] * xs + F[4] * ys + F[5];
        c = F[6] * xs + F[7] * ys + F[8];
    } else {
        a = F[0] * xs + F[3] * ys + F[6];
        b = F[1] * xs + F[4] * ys + F[7];
        c = F[2] * xs + F[5] * ys + F[8];
    }
    double n2 = a * a + b * b;
    return fabs(xt * a + yt * b + c) * rsqrt_f64(n2);
}

// epipolar_distance(cam1, person1, cam2, person2)[j] = [d1, d2] with F = cam1.F[cam2.cid]:
//   d1 = distance of x1 to the line F x2 (in image 1), d2 = distance of x2 to F^T x1 (in image 2).
PAM_HD void epi_pair_cv(const double* F12, double u1, double v1, double u2, double v2, double& d1, double& d2) {
    d1 = epi_dist_cv(F12, false, u2, v2, u1, v1);
    d2 = epi_dist_cv(F12, true, u1, v1, u2, v2);
}

// ------------------------------------------------------------------------------------------
// back-projected ray to 3-D point distance
// ------------------------------------------------------------------------------------------
// RK = R^-1 K^-1 (row-major 3x3), pos = camera centre, (u, v) pixel, X = 3-D point:
// || dir x (pos - X) || / || dir ||   (utils/matching.py:10-17 + utils/calculate.py:26-32; the
// reference normalises dir first and re-derives it as (pos + dir) - pos, a rounding-level detail).
PAM_HD double ray_point_distance(const double* RK, const double* pos, double u, double v, const double* X) {
    double a0 = RK[0] * u + RK[1] * v + RK[2];
    double a1 = RK[3] * u + RK[4] * v + RK[5];
    double a2 = RK[6] * u + RK[7] * v + RK[8];
    double b0 = pos[0] - X[0], b1 = pos[1] - X[1], b2 = pos[2] - X[2];
    double c0 = a1 * b2 - a2 * b1;
    double c1 = a2 * b0 - a0 * b2;
    double c2 = a0 * b1 - a1 * b0;
    return sqrt_f64((c0 * c0 + c1 * c1 + c2 * c2) * rcp_f64(a0 * a0 + a1 * a1 + a2 * a2));
}

// ------------------------------------------------------------------------------------------
// triangulation: weighted homogeneous DLT
// ------------------------------------------------------------------------------------------
// The reference stacks the 2k weighted unit rows of the k surviving views and takes the right
// singular vector of the smallest singular value (la.svd, utils/construction.py:109-113).  Here the
// system never exists in memory: rows are folded, as they are produced, into the 10 entries of the
// upper-triangular 4x4 factor R with A^T A = R^T R, and the singular vector is extracted from R.
//
//   fold      Gram matrix + Cholesky (10 FMA per row) while the pivots stay above 1e-6 of the
//             diagonal; otherwise (systems resting on views 2-3 frames old, weights e^{-lambda_t T}
//             down to 3e-7, or noise-free data) streaming Givens QR, which keeps the relative
//             accuracy of the tiny rows that a Gram matrix would lose
//   extract   inverse iteration with R (two triangular solves per step, converges like
//             (sigma4/sigma3)^2); if it has not settled in 8 steps, or a pivot is degenerate,
//             one-sided Jacobi SVD of R (unconditionally robust)
#if defined(PAM_COUNT_ITERS) && !defined(__CUDA_ARCH__)
static long long g_invit_steps = 0, g_invit_calls = 0;
#endif
struct DltAccum {
    double r00, r01, r02, r03, r11, r12, r13, r22, r23, r33;   // Gram entries, then R
    bool gram;

    PAM_HD void reset(bool use_gram) {
        r00 = r01 = r02 = r03 = r11 = r12 = r13 = r22 = r23 = r33 = 0.0;
        gram = use_gram;
    }

    PAM_HD static void givens(double& d, double& x, double& c, double& s) {
        double h2 = d * d + x * x;
        double inv = rsqrt_f64(h2);
        c = d * inv; s = x * inv;
        d = h2 * inv; x = 0.0;
    }
    PAM_HD void add_row(double x0, double x1, double x2, double x3) {
        if (gram) {
            r00 += x0 * x0; r01 += x0 * x1; r02 += x0 * x2; r03 += x0 * x3;
            r11 += x1 * x1; r12 += x1 * x2; r13 += x1 * x3;
            r22 += x2 * x2; r23 += x2 * x3;
            r33 += x3 * x3;
            return;
        }
        double c, s, t;
        if (x0 != 0.0) {
            givens(r00, x0, c, s);
            t = r01; r01 = c * t + s * x1; x1 = c * x1 - s * t;
            t = r02; r02 = c * t + s * x2; x2 = c * x2 - s * t;
            t = r03; r03 = c * t + s * x3; x3 = c * x3 - s * t;
        }
        if (x1 != 0.0) {
            givens(r11, x1, c, s);
            t = r12; r12 = c * t + s * x2; x2 = c * x2 - s * t;
            t = r13; r13 = c * t + s * x3; x3 = c * x3 - s * t;
        }
        if (x2 != 0.0) {
            givens(r22, x2, c, s);
            t = r23; r23 = c * t + s * x3; x3 = c * x3 - s * t;
        }
        if (x3 != 0.0) {
            givens(r33, x3, c, s);
        }
    }
    // the two DLT rows of one view (utils/construction.py:92-97):
    //   (u P2 - P0)/||.|| * w ,  (v P2 - P1)/||.|| * w      with P row-major 3x4
    PAM_HD void add_view(const double* P, double u, double v, double w) {
        double a0 = u * P[8] - P[0], a1 = u * P[9] - P[1], a2 = u * P[10] - P[2], a3 = u * P[11] - P[3];
        double sa = w * rsqrt_f64(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3);
        add_row(a0 * sa, a1 * sa, a2 * sa, a3 * sa);
        double b0 = v * P[8] - P[4], b1 = v * P[9] - P[5], b2 = v * P[10] - P[6], b3 = v * P[11] - P[7];
        double sb = w * rsqrt_f64(b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3);
        add_row(b0 * sb, b1 * sb, b2 * sb, b3 * sb);
    }

    // Gram -> R in place.  false when a pivot is too small for the squared system to be trusted.
    PAM_HD bool cholesky() {
        const double g11 = r11, g22 = r22, g33 = r33;
        if (!(r00 > 0.0)) return false;
        double i0 = rsqrt_f64(r00);
        r00 *= i0; r01 *= i0; r02 *= i0; r03 *= i0;
        double t11 = r11 - r01 * r01;
        if (!(t11 > 1e-6 * g11)) return false;
        double i1 = rsqrt_f64(t11);
        r11 = t11 * i1;
        r12 = (r12 - r01 * r02) * i1;
        r13 = (r13 - r01 * r03) * i1;
        double t22 = r22 - r02 * r02 - r12 * r12;
        if (!(t22 > 1e-6 * g22)) return false;
        double i2 = rsqrt_f64(t22);
        r22 = t22 * i2;
        r23 = (r23 - r02 * r03 - r12 * r13) * i2;
        double t33 = r33 - r03 * r03 - r13 * r13 - r23 * r23;
        if (!(t33 > 1e-13 * g33)) return false;
        r33 = t33 * rsqrt_f64(t33);
        return true;
    }

    // inverse iteration on R^T R; x = homogeneous solution (unit norm).  false = not converged.
    PAM_HD bool invit(double* x) const {
#if defined(PAM_COUNT_ITERS) && !defined(__CUDA_ARCH__)
        ++g_invit_calls;
#endif
        if (r00 == 0.0 || r11 == 0.0 || r22 == 0.0 || r33 == 0.0) return false;
        const double i0 = rcp_f64(r00), i1 = rcp_f64(r11), i2 = rcp_f64(r22), i3 = rcp_f64(r33);
        // start from R^-1 e4 (the direction R shrinks most when r33 is its smallest pivot)
        double x3 = 1.0;
        double x2 = -(r23 * x3) * i2;
        double x1 = -(r12 * x2 + r13 * x3) * i1;
        double x0 = -(r01 * x1 + r02 * x2 + r03 * x3) * i0;
        double inv = rsqrt_f64(x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3);
        x0 *= inv; x1 *= inv; x2 *= inv; x3 *= inv;
        bool ok = false;
        double e_prev = 0.0;
        PAM_NOUNROLL for (int it = 0; it < 8; ++it) {
            double y0 = x0 * i0;
            double y1 = (x1 - r01 * y0) * i1;
            double y2 = (x2 - r02 * y0 - r12 * y1) * i2;
            double y3 = (x3 - r03 * y0 - r13 * y1 - r23 * y2) * i3;
            double z3 = y3 * i3;
            double z2 = (y2 - r23 * z3) * i2;
            double z1 = (y1 - r12 * z2 - r13 * z3) * i1;
            double z0 = (y0 - r01 * z1 - r02 * z2 - r03 * z3) * i0;
            inv = rsqrt_f64(z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3);
            if (z0 * x0 + z1 * x1 + z2 * x2 + z3 * x3 < 0.0) inv = -inv;
            z0 *= inv; z1 *= inv; z2 *= inv; z3 *= inv;
            double e0 = z0 - x0, e1 = z1 - x1, e2 = z2 - x2, e3 = z3 - x3;
            x0 = z0; x1 = z1; x2 = z2; x3 = z3;
#if defined(PAM_COUNT_ITERS) && !defined(__CUDA_ARCH__)
            ++g_invit_steps;
#endif
            // The iteration converges linearly (ratio = (sigma_4 / sigma_3)^2, ~1e-5 for a triangulation with pixel
            // noise), so the step just taken, e, bounds the error of the previous iterate and e * (e / e_prev) that of
            // this one: stop when the step is below 1e-13, or when the predicted error is below 1e-15 (which spares
            // the iteration that would only confirm it).
            const double ee = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            if (ee <= 1e-26 || ee * ee <= 1e-30 * e_prev) { ok = true; break; }
            e_prev = ee;
        }
        x[0] = x0; x[1] = x1; x[2] = x2; x[3] = x3;
        return ok;
    }

    // One-sided (Hestenes) Jacobi on the columns of R: right singular vectors accumulate in V.
    // Columns are orthogonalised to |g_p . g_q| <= 1e-14 |g_p||g_q| (a relative criterion, so the
    // tiny singular values of stale-view systems are resolved as well as the large ones).
    PAM_HD_NOINLINE void jacobi(double* x) const {
        // rare fallback: kept compact (rolled loops, arrays in local memory) to spare the instruction cache
        double g[4][4] = {{r00, r01, r02, r03}, {0.0, r11, r12, r13}, {0.0, 0.0, r22, r23}, {0.0, 0.0, 0.0, r33}};
        double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
        for (int sweep = 0; sweep < 10; ++sweep) {
            bool rotated = false;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
            for (int pq = 0; pq < 6; ++pq) {
                const int p = (pq < 3) ? 0 : ((pq < 5) ? 1 : 2);
                const int q = (pq < 3) ? pq + 1 : ((pq < 5) ? pq - 1 : 3);
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int k = 0; k < 4; ++k) {
                    al += g[k][p] * g[k][p];
                    be += g[k][q] * g[k][q];
                    ga += g[k][p] * g[k][q];
                }
                if (ga * ga <= 1e-28 * (al * be)) continue;
                rotated = true;
                // tan of the rotation angle: t = sign(tau) 2 ga / (|tau| + sqrt(tau^2 + 4 ga^2)), tau = be - al
                const double tau = be - al, g2 = 2.0 * ga;
                const double h = sqrt_f64(tau * tau + g2 * g2);
                const double t = g2 * rcp_f64((tau < 0.0) ? (tau - h) : (tau + h));
                const double c = rsqrt_f64(1.0 + t * t), s = c * t;
                for (int k = 0; k < 4; ++k) {
                    const double gp = g[k][p], gq = g[k][q];
                    g[k][p] = c * gp - s * gq;
                    g[k][q] = s * gp + c * gq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - s * vq;
                    V[k][q] = s * vp + c * vq;
                }
            }
            if (!rotated) break;
        }
        double best = 0.0;
        int arg = 0;
        for (int q = 0; q < 4; ++q) {
            const double nq = g[0][q] * g[0][q] + g[1][q] * g[1][q] + g[2][q] * g[2][q] + g[3][q] * g[3][q];
            if (q == 0 || nq < best) { best = nq; arg = q; }
        }
        x[0] = V[0][arg]; x[1] = V[1][arg]; x[2] = V[2][arg]; x[3] = V[3][arg];
    }

    // homogeneous solution of a Gram-accumulated system: unit 4-vector with a positive last component (the
    // convention of the pair-wise triangulation below); false when the system is singular
    PAM_HD bool solve_homog(double* x) {
        if (gram && !cholesky()) return false;
        if (!invit(x)) jacobi(x);
        if (x[3] < 0.0) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; x[3] = -x[3]; }
        return true;
    }

    // de-homogenised joint.  `path` (optional) reports which extractor produced it (tests).
    PAM_HD void solve(double* X, int* path = nullptr) {
        double x[4];
        int how = 0;
        if (gram && !cholesky()) {
            // a squared system this close to singular is not trusted: signal the caller to refold
            X[0] = X[1] = X[2] = 0.0;
            if (path) *path = -1;
            return;
        }
        if (!invit(x)) { jacobi(x); how = 1; }
        double ix = rcp_f64(x[3]);
        X[0] = x[0] * ix; X[1] = x[1] * ix; X[2] = x[2] * ix;
        if (path) *path = how;
    }
};

// ------------------------------------------------------------------------------------------
// rectangular linear sum assignment (scipy rectangular_lsap.cpp / Crouse 2016)
// ------------------------------------------------------------------------------------------
// cost(i, j) for i < nr, j < nc (minimised).  col4row[i] = assigned column or -1.  When nc < nr
// the transposed problem is solved, exactly as scipy does.  Returns 0, or -1 if infeasible.
template <int MAXN, class CostFn>
PAM_HD_NOINLINE int lsap_solve(int nr0, int nc0, CostFn cost, int* col4row_out) {
    PAM_NOUNROLL for (int i = 0; i < nr0; ++i) col4row_out[i] = -1;
    if (nr0 == 0 || nc0 == 0) return 0;
    const bool transpose = nc0 < nr0;
    const int nr = transpose ? nc0 : nr0, nc = transpose ? nr0 : nc0;
    double u[MAXN], v[MAXN], sp[MAXN];
    int path[MAXN], col4row[MAXN], row4col[MAXN], remaining[MAXN];
    uint64_t SR, SC;
    PAM_NOUNROLL for (int i = 0; i < nr; ++i) { u[i] = 0.0; col4row[i] = -1; }
    PAM_NOUNROLL for (int j = 0; j < nc; ++j) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    const double INF = HUGE_VAL;
    PAM_NOUNROLL for (int cur = 0; cur < nr; ++cur) {
        double minVal = 0.0;
        int i = cur;
        int num_remaining = nc;
        PAM_NOUNROLL for (int it = 0; it < nc; ++it) { remaining[it] = nc - it - 1; sp[it] = INF; }
        SR = 0ull; SC = 0ull;
        int sink = -1;
        PAM_NOUNROLL while (sink == -1) {
            int index = -1;
            double lowest = INF;
            SR |= (1ull << i);
            PAM_NOUNROLL for (int it = 0; it < num_remaining; ++it) {
                int j = remaining[it];
                double cij = transpose ? cost(j, i) : cost(i, j);
                double r = minVal + cij - u[i] - v[j];
                if (r < sp[j]) { path[j] = i; sp[j] = r; }
                if (sp[j] < lowest || (sp[j] == lowest && row4col[j] == -1)) { lowest = sp[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INF) return -1;
            int j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC |= (1ull << j);
            remaining[index] = remaining[--num_remaining];
        }
        u[cur] += minVal;
        PAM_NOUNROLL for (int k = 0; k < nr; ++k)
            if (((SR >> k) & 1ull) && k != cur) u[k] += minVal - sp[col4row[k]];
        PAM_NOUNROLL for (int j = 0; j < nc; ++j)
            if ((SC >> j) & 1ull) v[j] -= minVal - sp[j];
        int j = sink;
        PAM_NOUNROLL while (true) {
            int k = path[j];
            row4col[j] = k;
            int tmp = col4row[k]; col4row[k] = j; j = tmp;
            if (k == cur) break;
        }
    }
    if (!transpose) {
        PAM_NOUNROLL for (int i = 0; i < nr; ++i) col4row_out[i] = col4row[i];
    } else {
        // col4row maps (original column) -> (original row)
        PAM_NOUNROLL for (int c = 0; c < nr; ++c) col4row_out[col4row[c]] = c;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// temporal smoothing
// ------------------------------------------------------------------------------------------
// scipy 'reflect' (half-sample symmetric) extension: d c b a | a b c d | d c b a
PAM_HD int reflect_index(int i, int n) {
    PAM_NOUNROLL while (i < 0 || i >= n) {
        if (i < 0) i = -1 - i;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i;
}

// Gaussian weights of scipy.ndimage._gaussian_kernel1d, order 0: w[k] for |offset| = k.
inline int gaussian_weights(double sigma, double* w /* PAM_MAX_RADIUS+1 */) {
    int radius = (int)(4.0 * sigma + 0.5);
    if (radius > PAM_MAX_RADIUS) return -1;
    double phi[2 * PAM_MAX_RADIUS + 1];
    double sigma2 = sigma * sigma;
    for (int x = -radius; x <= radius; ++x) phi[x + radius] = exp(-0.5 / sigma2 * (double)(x * x));
    // numpy pairwise order for phi.sum()
    double s = np_sum(phi, 2 * radius + 1);
    for (int k = 0; k <= radius; ++k) w[k] = phi[radius + k] / s;
    return radius;
}

}  // namespace pam
