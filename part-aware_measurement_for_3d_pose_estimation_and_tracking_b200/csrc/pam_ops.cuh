// pam_ops.cuh -- stateless batched kernels: one kernel per reference function of SURVEY.md
// section 8a that is callable on its own (the drop-in modules matching / construction / calculate
// bind to these through the C ABI).  All arithmetic is FP64 on the reference's float32 camera
// constants; any number of cameras up to PAM_OPS_MAX_V.
//
//   k_project_points      Camera.projectPoints_parallel                 ivclabpose.py:91-98
//   k_assoc_affinity      association block of tracking()               tracking/IterativeTracker.py:137-149
//   k_assign              scipy linear_sum_assignment (batched)         tracking/IterativeTracker.py:79,150
//   k_epipolar_pairs      epipolar_affinity_parallel                    utils/matching.py:115-151
//   k_epipolar_allpairs   epipolar_affinity (+ epipolar_distance)       utils/matching.py:50-113
//   k_view_filter         Greedy_matching, modes 'update' / 'init'      utils/matching.py:243-295
//   k_triangulate         SVD_pose_kernel_jf / _parallel / SVD_pose_kernel   utils/construction.py:64-131
//   k_ray_distance        back_project_ray + line2point_distance_3D     utils/matching.py:10-17, utils/calculate.py:26-32
//   k_epipolar_distance   epipolar_distance (pairwise, both directions) utils/matching.py:50-91
#pragma once
#include "pam_core.h"

#define PAM_OPS_MAX_V 32
#define PAM_OPS_MAX_N 64   // assignment problems up to 64 x 64

namespace pam {

__device__ __forceinline__ void load9(const float* __restrict__ src, double* dst) {
#pragma unroll
    for (int k = 0; k < 9; ++k) dst[k] = (double)__ldg(src + k);
}
__device__ __forceinline__ void load12(const float* __restrict__ src, double* dst) {
#pragma unroll
    for (int k = 0; k < 12; ++k) dst[k] = (double)__ldg(src + k);
}

// ---- a2: (n, J, 3) world joints -> (V, n, J, 2) pixels (v, u) ------------------------------------
__global__ void k_project_points(const float* __restrict__ P, int V, const double* __restrict__ X, int N,
                                 double* __restrict__ out) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= V * N) return;
    const int cam = it / N, p = it - cam * N;
    double Pm[12];
    load12(P + cam * 12, Pm);
    const double x = X[p * 3], y = X[p * 3 + 1], z = X[p * 3 + 2];
    const double a = Pm[0] * x + Pm[1] * y + Pm[2] * z + Pm[3];
    const double b = Pm[4] * x + Pm[5] * y + Pm[6] * z + Pm[7];
    const double w = Pm[8] * x + Pm[9] * y + Pm[10] * z + Pm[11];
    out[(int64_t)it * 2 + 0] = b / w;
    out[(int64_t)it * 2 + 1] = a / w;
}

// ---- a1: camera ingest (ivclabpose.py:35-46,162-181): K, RT -> RK_INV, centre, V x V fundamental tensor ----
// One thread per ordered camera pair.  Float32 with fused multiply-adds in the association order of
// the reference's torch expression  K_a^-T (R_a R_b^T) K_b^T [K_b R_b R_a^T (t_a - R_a R_b^T t_b)]_x ;
// the inverse is the adjugate form.  The reference's LAPACK/MKL kernels round differently in the
// last bits, so this op matches the host ingest (camera.fundamental_tensor) to float32 rounding, not
// bit for bit; the tracker's default ingest stays on the host for that reason.
__device__ __forceinline__ void mm33f(const float* A, const float* B, float* C) {      // C = A B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = fmaf(A[i * 3 + 2], B[6 + j], fmaf(A[i * 3 + 1], B[3 + j], A[i * 3] * B[j]));
}
__device__ __forceinline__ void mm33f_bt(const float* A, const float* B, float* C) {   // C = A B^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = fmaf(A[i * 3 + 2], B[j * 3 + 2], fmaf(A[i * 3 + 1], B[j * 3 + 1], A[i * 3] * B[j * 3]));
}
__device__ __forceinline__ void mv33f(const float* A, const float* x, float* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = fmaf(A[i * 3 + 2], x[2], fmaf(A[i * 3 + 1], x[1], A[i * 3] * x[0]));
}
template <typename T>
__device__ __forceinline__ void inv33(const T* A, T* I) {
    const T c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const T r = (T)1 / (A[0] * c0 + A[1] * c1 + A[2] * c2);
    I[0] = c0 * r; I[1] = (A[2] * A[7] - A[1] * A[8]) * r; I[2] = (A[1] * A[5] - A[2] * A[4]) * r;
    I[3] = c1 * r; I[4] = (A[0] * A[8] - A[2] * A[6]) * r; I[5] = (A[2] * A[3] - A[0] * A[5]) * r;
    I[6] = c2 * r; I[7] = (A[1] * A[6] - A[0] * A[7]) * r; I[8] = (A[0] * A[4] - A[1] * A[3]) * r;
}
__global__ void k_camera_ingest(int V, const float* __restrict__ K, const float* __restrict__ RT,
                                float* __restrict__ RKinv, double* __restrict__ pos, float* __restrict__ F) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= V * V) return;
    const int a = it / V, b = it - a * V;
    float Ka[9], Kb[9], Ra[9], Rb[9], ta[3], tb[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { Ka[k] = K[a * 9 + k]; Kb[k] = K[b * 9 + k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { Ra[i * 3 + j] = RT[a * 12 + i * 4 + j]; Rb[i * 3 + j] = RT[b * 12 + i * 4 + j]; }
        ta[i] = RT[a * 12 + i * 4 + 3]; tb[i] = RT[b * 12 + i * 4 + 3];
    }
    float Kai[9], KaiT[9], Rab[9], M1[9], M2[9], KR[9], KRR[9], v[3], e[3];
    inv33<float>(Ka, Kai);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) KaiT[i * 3 + j] = Kai[j * 3 + i];
    mm33f_bt(Ra, Rb, Rab);                 // R_a R_b^T
    mm33f(KaiT, Rab, M1);
    mm33f_bt(M1, Kb, M2);                  // ... K_b^T
    mm33f(Kb, Rb, KR);
    mm33f_bt(KR, Ra, KRR);                 // K_b R_b R_a^T
    mv33f(Rab, tb, v);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = ta[i] - v[i];
    mv33f(KRR, v, e);
    const float ex[9] = {0.f, -e[2], e[1], e[2], 0.f, -e[0], -e[1], e[0], 0.f};
    float Fm[9];
    mm33f(M2, ex, Fm);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) sum += Fm[k];
    if (sum == 0.f)                        // "to avoid nan", ivclabpose.py:176-177
#pragma unroll
        for (int k = 0; k < 9; ++k) Fm[k] += 1e-12f;
#pragma unroll
    for (int k = 0; k < 9; ++k) F[(int64_t)it * 9 + k] = Fm[k];
    if (a != b) return;
    // Camera.__init__: RK_INV = R^-1 K^-1 (float32); centre = -R^-1 t from the float64 4 x 4 inverse
    float Rai[9], RK[9];
    inv33<float>(Ra, Rai);
    mm33f(Rai, Kai, RK);
#pragma unroll
    for (int k = 0; k < 9; ++k) RKinv[a * 9 + k] = RK[k];
    double Rd[9], Rdi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rd[k] = (double)Ra[k];
    inv33<double>(Rd, Rdi);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        pos[a * 3 + i] = -(Rdi[i * 3] * (double)ta[0] + Rdi[i * 3 + 1] * (double)ta[1] + Rdi[i * 3 + 2] * (double)ta[2]);
}

// ---- a3: affinity of every (camera, track, detection) ---------------------------------------------
// tracks X [n][J][3], dt [n]; dets [V][mmax][J][3] f64 (v,u,conf), counts [V]; aff [V][n][mmax]
__global__ void k_assoc_affinity(const float* __restrict__ P, int V, const double* __restrict__ X,
                                 const int* __restrict__ dt, const double* __restrict__ dets,
                                 const int* __restrict__ counts, int n, int mmax, int J, double alpha2d,
                                 double lambda_a, int min_valid, double* __restrict__ aff) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= V * n * mmax) return;
    const int cam = it / (n * mmax), rem = it - cam * (n * mmax), i = rem / mmax, d = rem - i * mmax;
    if (d >= counts[cam]) { aff[it] = 0.0; return; }
    double Pm[12];
    load12(P + cam * 12, Pm);
    const double* q = dets + ((int64_t)(cam * mmax + d) * J) * 3;
    const double* Xi = X + (int64_t)i * J * 3;
    const double denom = alpha2d * (double)dt[i];
    double sum = 0.0;
    int cnt = 0;
    for (int j = 0; j < J; ++j) {
        const double x = Xi[j * 3], y = Xi[j * 3 + 1], z = Xi[j * 3 + 2];
        const double a = Pm[0] * x + Pm[1] * y + Pm[2] * z + Pm[3];
        const double b = Pm[4] * x + Pm[5] * y + Pm[6] * z + Pm[7];
        const double w = Pm[8] * x + Pm[9] * y + Pm[10] * z + Pm[11];
        const double dv = b / w - q[j * 3 + 0], du = a / w - q[j * 3 + 1];
        const double cj = 1.0 - sqrt(dv * dv + du * du) / denom;
        if (cj > 0.0) { sum += cj; ++cnt; }
    }
    double r = (cnt > min_valid) ? sum / (double)cnt : 0.0;
    r = r / exp(lambda_a * (double)dt[i]);
    if (r != r) r = 0.0;
    aff[it] = r;
}

// ---- a4: batched rectangular assignment; one thread per problem ----------------------------------
__global__ void k_assign(const double* __restrict__ cost, int B, int nr, int nc, int maximize, int* __restrict__ col4row) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* C = cost + (int64_t)b * nr * nc;
    const double sgn = maximize ? -1.0 : 1.0;
    int out[PAM_OPS_MAX_N];
    lsap_solve<PAM_OPS_MAX_N>(nr, nc, [&](int i, int j) { return sgn * C[i * nc + j]; }, out);
    for (int i = 0; i < nr; ++i) col4row[(int64_t)b * nr + i] = out[i];
}

// ---- a4, large problems: one WARP per assignment problem ----------------------------------------------
// Same shortest-augmenting-path algorithm (Crouse 2016) as lsap_solve, with the scan over the columns
// done by the 32 lanes in parallel: lane l owns columns l and l + 32 (dual v, shortest-path cost, path,
// row4col) and rows l and l + 32 (dual u, col4row); the minimum over the unscanned columns is a warp
// shuffle reduction.  Among equal minima an unassigned column is preferred (as scipy does), then the
// lowest column index (scipy: scan order of its `remaining` list), so with exact ties a different but
// equally optimal assignment may be returned.  nr, nc <= 64.
__device__ __forceinline__ double warp_pick(double a0, double a1, int lane, int k) {
    // value held by lane (k & 31) in register (k >> 5)
    const double x0 = __shfl_sync(0xffffffffu, a0, k & 31), x1 = __shfl_sync(0xffffffffu, a1, k & 31);
    return (k >> 5) ? x1 : x0;
}
__device__ __forceinline__ int warp_pick(int a0, int a1, int lane, int k) {
    const int x0 = __shfl_sync(0xffffffffu, a0, k & 31), x1 = __shfl_sync(0xffffffffu, a1, k & 31);
    return (k >> 5) ? x1 : x0;
}
__global__ void __launch_bounds__(128)
k_assign_warp(const double* __restrict__ cost, int B, int nr0, int nc0, int maximize, int* __restrict__ col4row_out) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const double* C = cost + (int64_t)b * nr0 * nc0;
    const double sgn = maximize ? -1.0 : 1.0;
    const bool tr = nc0 < nr0;
    const int nr = tr ? nc0 : nr0, nc = tr ? nr0 : nc0;
    auto cst = [&](int i, int j) { return sgn * (tr ? C[j * nc0 + i] : C[i * nc0 + j]); };
    const double INF = HUGE_VAL;
    double u[2] = {0.0, 0.0}, v[2] = {0.0, 0.0}, sp[2];
    int c4r[2] = {-1, -1}, r4c[2] = {-1, -1}, path[2] = {-1, -1};
    for (int cur = 0; cur < nr; ++cur) {
        double minVal = 0.0;
        int i = cur, sink = -1;
        uint32_t sc = 0u;                       // bit k: my column lane + 32 k is scanned
        uint32_t sr0 = 0u, sr1 = 0u;            // rows in SR (replicated on every lane)
        sp[0] = sp[1] = INF;
        while (sink < 0) {
            if (i < 32) sr0 |= 1u << i; else sr1 |= 1u << (i - 32);
            const double ui = warp_pick(u[0], u[1], lane, i);
            double best = INF;
            int bestj = -1, bestfree = 0;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = lane + 32 * k;
                if (j >= nc || ((sc >> k) & 1u)) continue;
                const double r = minVal + cst(i, j) - ui - v[k];
                if (r < sp[k]) { path[k] = i; sp[k] = r; }
                const int fr = (r4c[k] == -1) ? 1 : 0;
                if (sp[k] < best || (sp[k] == best && fr > bestfree)) { best = sp[k]; bestj = j; bestfree = fr; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bestj, off), of = __shfl_xor_sync(0xffffffffu, bestfree, off);
                const bool take = (oj >= 0) && (bestj < 0 || ob < best || (ob == best && (of > bestfree || (of == bestfree && oj < bestj))));
                if (take) { best = ob; bestj = oj; bestfree = of; }
            }
            minVal = best;
            if (bestj < 0 || minVal == INF) { sink = -2; break; }      // infeasible
            if ((bestj & 31) == lane) sc |= 1u << (bestj >> 5);
            if (bestfree) sink = bestj;
            else i = warp_pick(r4c[0], r4c[1], lane, bestj);
        }
        if (sink == -2) break;
        // dual updates
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int r = lane + 32 * k;
            // every lane takes part in the shuffles; only rows in SR (other than cur) are updated
            const int col = (r < nr && c4r[k] >= 0) ? c4r[k] : 0;
            const double spc = warp_pick(sp[0], sp[1], lane, col);
            const bool in_sr = r < nr && (((k ? sr1 : sr0) >> lane) & 1u);
            if (in_sr && r != cur) u[k] += minVal - spc;
            if (r == cur) u[k] += minVal;
            if ((sc >> k) & 1u) v[k] -= minVal - sp[k];
        }
        // augment along the path
        int j = sink;
        while (true) {
            const int pi = warp_pick(path[0], path[1], lane, j);
            if ((j & 31) == lane) r4c[j >> 5] = pi;
            const int old = warp_pick(c4r[0], c4r[1], lane, pi);
            if ((pi & 31) == lane) c4r[pi >> 5] = j;
            j = old;
            if (pi == cur) break;
        }
    }
    // rows of the ORIGINAL problem
    if (!tr) {
#pragma unroll
        for (int k = 0; k < 2; ++k) { const int r = lane + 32 * k; if (r < nr0) col4row_out[(int64_t)b * nr0 + r] = c4r[k]; }
    } else {
        // c4r maps original column -> original row; unassigned original rows stay -1 (pre-set by the host)
#pragma unroll
        for (int k = 0; k < 2; ++k) { const int r = lane + 32 * k; if (r < nr && c4r[k] >= 0) col4row_out[(int64_t)b * nr0 + c4r[k]] = r; }
    }
}

// ---- a6: symmetric point-to-epipolar-line distances, float64 form ----------------------------------
// pose [M][J][3] (v,u,conf), cam [M]; D [M][M][J], mean [M][M]; same-camera pairs use F = 0.
__global__ void k_epipolar_pairs(const float* __restrict__ F, int V, const double* __restrict__ pose,
                                 const int* __restrict__ cam, int M, int J, double* __restrict__ D,
                                 double* __restrict__ mean) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= M * M) return;
    const int a = it / M, b = it - a * M;
    if (a > b) return;                        // the (a, b) thread writes both triangles
    const int ca = cam[a], cb = cam[b];
    double Fab[9], Fba[9];
    if (ca == cb) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { Fab[k] = 0.0; Fba[k] = 0.0; }
    } else {
        load9(F + ((int64_t)ca * V + cb) * 9, Fab);
        load9(F + ((int64_t)cb * V + ca) * 9, Fba);
    }
    const double* pa = pose + (int64_t)a * J * 3;
    const double* pb = pose + (int64_t)b * J * 3;
    double vals[PAM_MAX_J];
    for (int j = 0; j < J; ++j) {
        const double va = pa[j * 3], ua = pa[j * 3 + 1], vb = pb[j * 3], ub = pb[j * 3 + 1];
        const double dab = epi_dist_f64(Fab, ua, va, ub, vb);
        const double dba = epi_dist_f64(Fba, ub, vb, ua, va);
        const double d = (dab + dba) / 2.0;
        vals[j] = d;
        D[((int64_t)a * M + b) * J + j] = d;
        D[((int64_t)b * M + a) * J + j] = d;
    }
    const double m = np_sum(vals, J) / (double)J;
    mean[(int64_t)a * M + b] = m;
    mean[(int64_t)b * M + a] = m;
}

// ---- a7: all-pairs form with float32 stores (dense crowd) ------------------------------------------
// One CTA = a TILE x TILE block of detection pairs (upper-triangular tiles only; results are
// mirrored).  Both pose tiles are staged in shared memory as (u, v) pairs; each thread walks the
// joints of its pairs in FP64.  aff [M][M] f32 (25 between detections of one camera, 0 on the
// diagonal).  The optional per-joint tensor D [M][M][J] f32 (299 MB at M = 1984) is staged per tile
// in shared memory and written out as contiguous TILE*J-float runs in BOTH orientations, so the
// mirrored half is coalesced too.
#define PAM_AP_TILE 32
__global__ void __launch_bounds__(256)
k_epipolar_allpairs(const float* __restrict__ F, int V, const double* __restrict__ pose, const int* __restrict__ cam,
                    int M, int J, float* __restrict__ aff, float* __restrict__ D) {
    extern __shared__ double tile[];           // [2][TILE][J][2] doubles, then (if D) [TILE][TILE][J] floats
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;
    double* A = tile;
    double* Bt = tile + PAM_AP_TILE * J * 2;
    float* Ds = (float*)(tile + 2 * PAM_AP_TILE * J * 2);
    for (int e = threadIdx.x; e < PAM_AP_TILE * J; e += blockDim.x) {
        const int r = e / J, j = e - r * J;
        const int ia = ti * PAM_AP_TILE + r, ib = tj * PAM_AP_TILE + r;
        if (ia < M) { A[e * 2] = pose[((int64_t)ia * J + j) * 3 + 1]; A[e * 2 + 1] = pose[((int64_t)ia * J + j) * 3]; }
        if (ib < M) { Bt[e * 2] = pose[((int64_t)ib * J + j) * 3 + 1]; Bt[e * 2 + 1] = pose[((int64_t)ib * J + j) * 3]; }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < PAM_AP_TILE * PAM_AP_TILE; p += blockDim.x) {
        const int r = p / PAM_AP_TILE, cidx = p - r * PAM_AP_TILE;
        const int i = ti * PAM_AP_TILE + r, j2 = tj * PAM_AP_TILE + cidx;
        float* ds = D ? Ds + (int64_t)p * J : nullptr;
        if (i >= M || j2 >= M) continue;
        const int ci = cam[i], cj = cam[j2];
        if (i == j2 || ci == cj || (ti == tj && j2 < i)) {
            // diagonal: 0; same camera: 25 / zeros (matching.py:97-102); lower half of a diagonal tile:
            // filled by its mirror below
            if (i == j2) aff[(int64_t)i * M + i] = 0.0f;
            else if (ci == cj) aff[(int64_t)i * M + j2] = 25.0f;
            if (ds && !(ti == tj && j2 < i && ci != cj)) for (int j = 0; j < J; ++j) ds[j] = 0.0f;
            if (ci == cj && i != j2 && ti != tj) aff[(int64_t)j2 * M + i] = 25.0f;
            continue;
        }
        double Fm[9];
        load9(F + ((int64_t)ci * V + cj) * 9, Fm);
        const double* xa = A + (int64_t)r * J * 2;
        const double* xb = Bt + (int64_t)cidx * J * 2;
        NpSumStream<double> acc;
        acc.begin(J);
        for (int j = 0; j < J; ++j) {
            double d1, d2;
            epi_pair_cv(Fm, xa[j * 2], xa[j * 2 + 1], xb[j * 2], xb[j * 2 + 1], d1, d2);
            const double sym = (d1 + d2) / 2.0;
            acc.push(sym);
            if (ds) ds[j] = (float)sym;
        }
        const float m = (float)(acc.total() / (double)J);
        aff[(int64_t)i * M + j2] = m;
        aff[(int64_t)j2 * M + i] = m;
        if (ds && ti == tj) {                  // mirror inside a diagonal tile
            float* dm = Ds + (int64_t)(cidx * PAM_AP_TILE + r) * J;
            for (int j = 0; j < J; ++j) dm[j] = ds[j];
        }
    }
    if (!D) return;
    __syncthreads();
    // row i of the tile = TILE*J contiguous floats of D[i][tj*TILE ...]; and, for off-diagonal tiles,
    // column c of the tile = TILE*J contiguous floats of D[j2][ti*TILE ...]
    const int run = PAM_AP_TILE * J;
    const int cols = min(PAM_AP_TILE, M - tj * PAM_AP_TILE), rows = min(PAM_AP_TILE, M - ti * PAM_AP_TILE);
    for (int e = threadIdx.x; e < PAM_AP_TILE * run; e += blockDim.x) {
        const int r = e / run, k = e - r * run;          // k = cidx * J + j
        if (r < rows && k < cols * J)
            D[((int64_t)(ti * PAM_AP_TILE + r) * M + tj * PAM_AP_TILE) * J + k] = Ds[(int64_t)r * run + k];
    }
    if (ti != tj) {
        for (int e = threadIdx.x; e < PAM_AP_TILE * run; e += blockDim.x) {
            const int cidx = e / run, k = e - cidx * run;   // k = r * J + j
            const int r = k / J, j = k - r * J;
            if (cidx < cols && r < rows)
                D[((int64_t)(tj * PAM_AP_TILE + cidx) * M + ti * PAM_AP_TILE) * J + k] = Ds[(int64_t)(r * PAM_AP_TILE + cidx) * J + j];
        }
    }
}

// ---- a8 / a17: part-aware view filter on given affinity matrices -----------------------------------
// B problems of Vt views: A [B][Vt][Vt] (f64, or f32 values widened by the caller for 'init'),
// uv [B][Vt][2] (u, v), cam [Vt], next [B][3]  ->  keep [B][Vt] (0/1).  mode 0 = update, 1 = init.
__global__ void k_view_filter(const float* __restrict__ RKinv, const double* __restrict__ pos,
                              const double* __restrict__ A, const float* __restrict__ Af, const double* __restrict__ uv,
                              const int* __restrict__ cam, const double* __restrict__ next, int B, int Vt, int mode,
                              unsigned char* __restrict__ keep) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    uint32_t alive = (Vt >= 32) ? 0xffffffffu : ((1u << Vt) - 1u);
    if (mode == 0) {
        const double* Ab = A + (int64_t)b * Vt * Vt;
        double rd[PAM_OPS_MAX_V];
        for (int a = 0; a < Vt; ++a) rd[a] = 0.0;
        for (int r = 0; r < Vt; ++r)
            for (int c2 = r; c2 < Vt; ++c2) {
                if (!(Ab[r * Vt + c2] < 0.0)) continue;
                if (!((alive >> r) & 1u) || !((alive >> c2) & 1u)) continue;
                for (int s = 0; s < 2; ++s) {
                    const int k = s ? c2 : r;
                    if (rd[k] == 0.0) {
                        double RK[9], pc[3] = {pos[cam[k] * 3], pos[cam[k] * 3 + 1], pos[cam[k] * 3 + 2]};
                        load9(RKinv + cam[k] * 9, RK);
                        rd[k] = ray_point_distance(RK, pc, uv[((int64_t)b * Vt + k) * 2], uv[((int64_t)b * Vt + k) * 2 + 1],
                                                   next + (int64_t)b * 3);
                    }
                }
                if (rd[r] > rd[c2]) alive &= ~(1u << r); else alive &= ~(1u << c2);
            }
    } else {
        const float* Ab = Af + (int64_t)b * Vt * Vt;
        for (int r = 0; r < Vt; ++r)
            for (int c2 = r; c2 < Vt; ++c2) {
                if (!(Ab[r * Vt + c2] < 0.0f)) continue;
                if (!((alive >> r) & 1u) || !((alive >> c2) & 1u)) continue;
                const float s1 = np_sum(Ab + r * Vt, Vt), s2 = np_sum(Ab + c2 * Vt, Vt);
                if (s1 > s2) alive &= ~(1u << c2); else alive &= ~(1u << r);
            }
    }
    for (int a = 0; a < Vt; ++a) keep[(int64_t)b * Vt + a] = (alive >> a) & 1u;
}

// ---- a10 / a11: weighted DLT of B x J joints over Vt views -------------------------------------------
// pose [B][Vt][J][3] (v,u,conf), cam [B][Vt], w [B][Vt] = exp(-lambda_t T), keep [B][J][Vt] or null,
// next [B][J][3] or null (used when < 2 views survive; NaN if null)  ->  out [B][J][3]
// LANES threads cooperate on one joint: with all-fresh views the Gram matrix is additive, so each lane
// folds every LANES-th view and the 10 entries are summed with warp shuffles (dense rigs: 31 views);
// systems with stale views (weights < 1) are folded with Givens rotations by the first lane.
template <int LANES>
__global__ void k_triangulate(const float* __restrict__ P, const double* __restrict__ pose, const int* __restrict__ cam,
                              const double* __restrict__ w, const unsigned char* __restrict__ keep,
                              const double* __restrict__ next, int B, int Vt, int J, double* __restrict__ out) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int it = gt / LANES, lane = gt % LANES;
    const bool active = it < B * J;
    const int b = active ? it / J : 0, j = active ? it - b * J : 0;
    const unsigned char* kp = (keep && active) ? keep + (int64_t)it * Vt : nullptr;
    int nv = 0;
    bool fresh = true;
    if (active)
        for (int a = 0; a < Vt; ++a) {
            if (kp && !kp[a]) continue;
            ++nv;
            if (w[(int64_t)b * Vt + a] != 1.0) fresh = false;
        }
    double* o = out + (int64_t)it * 3;
    if (active && nv < 2 && lane == 0) {
        if (next) { o[0] = next[(int64_t)it * 3]; o[1] = next[(int64_t)it * 3 + 1]; o[2] = next[(int64_t)it * 3 + 2]; }
        else { o[0] = o[1] = o[2] = nan(""); }
    }
    const bool work = active && nv >= 2;
    DltAccum acc;
    int path = -1;
    {
        acc.reset(true);
        if (work)
            for (int a = lane; a < Vt; a += LANES) {
                if (kp && !kp[a]) continue;
                double Pm[12];
                load12(P + cam[(int64_t)b * Vt + a] * 12, Pm);
                const double* q = pose + (((int64_t)b * Vt + a) * J + j) * 3;
                acc.add_view(Pm, q[1], q[0], w[(int64_t)b * Vt + a]);
            }
        if (LANES > 1) {
#pragma unroll
            for (int off = LANES / 2; off > 0; off >>= 1) {
                acc.r00 += __shfl_down_sync(0xffffffffu, acc.r00, off, LANES);
                acc.r01 += __shfl_down_sync(0xffffffffu, acc.r01, off, LANES);
                acc.r02 += __shfl_down_sync(0xffffffffu, acc.r02, off, LANES);
                acc.r03 += __shfl_down_sync(0xffffffffu, acc.r03, off, LANES);
                acc.r11 += __shfl_down_sync(0xffffffffu, acc.r11, off, LANES);
                acc.r12 += __shfl_down_sync(0xffffffffu, acc.r12, off, LANES);
                acc.r13 += __shfl_down_sync(0xffffffffu, acc.r13, off, LANES);
                acc.r22 += __shfl_down_sync(0xffffffffu, acc.r22, off, LANES);
                acc.r23 += __shfl_down_sync(0xffffffffu, acc.r23, off, LANES);
                acc.r33 += __shfl_down_sync(0xffffffffu, acc.r33, off, LANES);
            }
        }
        if (work && lane == 0) acc.solve(o, &path);
    }
    if (work && lane == 0 && path < 0) {
        acc.reset(false);
        for (int a = 0; a < Vt; ++a) {
            if (kp && !kp[a]) continue;
            double Pm[12];
            load12(P + cam[(int64_t)b * Vt + a] * 12, Pm);
            const double* q = pose + (((int64_t)b * Vt + a) * J + j) * 3;
            acc.add_view(Pm, q[1], q[0], w[(int64_t)b * Vt + a]);
        }
        acc.solve(o, &path);
    }
}

// ---- a9: distance of 3-D points to back-projected pixel rays ---------------------------------------
// uv [n][2] (u, v), X [n][3] -> dist [n], dirs [n][3] (unit) for camera `cam`
__global__ void k_ray_distance(const float* __restrict__ RKinv, const double* __restrict__ pos, int cam,
                               const double* __restrict__ uv, const double* __restrict__ X, int n,
                               double* __restrict__ dist, double* __restrict__ dirs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double RK[9], pc[3] = {pos[cam * 3], pos[cam * 3 + 1], pos[cam * 3 + 2]};
    load9(RKinv + cam * 9, RK);
    const double u = uv[i * 2], v = uv[i * 2 + 1];
    if (dirs) {
        double d0 = RK[0] * u + RK[1] * v + RK[2], d1 = RK[3] * u + RK[4] * v + RK[5], d2 = RK[6] * u + RK[7] * v + RK[8];
        const double inv = 1.0 / sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        dirs[i * 3] = d0 * inv; dirs[i * 3 + 1] = d1 * inv; dirs[i * 3 + 2] = d2 * inv;
    }
    if (dist && X) dist[i] = ray_point_distance(RK, pc, u, v, X + (int64_t)i * 3);
}

// ---- a7 (pairwise form): epipolar_distance(cam1, person1, cam2, person2) -> (n, 2) -----------------
// B pairs of poses: p1 [B][J][3], p2 [B][J][3] with cameras c1[B], c2[B] -> out [B][J][2]
__global__ void k_epipolar_distance(const float* __restrict__ F, int V, const double* __restrict__ p1,
                                    const double* __restrict__ p2, const int* __restrict__ c1, const int* __restrict__ c2,
                                    int B, int J, double* __restrict__ out) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= B * J) return;
    const int b = it / J;
    double Fm[9];
    load9(F + ((int64_t)c1[b] * V + c2[b]) * 9, Fm);
    double d1, d2;
    epi_pair_cv(Fm, p1[(int64_t)it * 3 + 1], p1[(int64_t)it * 3], p2[(int64_t)it * 3 + 1], p2[(int64_t)it * 3], d1, d2);
    out[(int64_t)it * 2] = d1;
    out[(int64_t)it * 2 + 1] = d2;
}

// ---- a16: get_believe (utils/calculate.py:8-14), batched: mean confidence over joints with conf >= 0 ----
__global__ void k_mean_confidence(const double* __restrict__ pose, int B, int J, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double kept[PAM_MAX_J];
    int nk = 0;
    for (int j = 0; j < J; ++j) {
        const double w = pose[((int64_t)b * J + j) * 3 + 2];
        if (w >= 0.0) kept[nk++] = w;
    }
    out[b] = np_sum(kept, nk) / (double)nk;      // 0/0 -> NaN like np.mean([])
}

// ---- a16: Hypothesis.calculate_cost (tracking/hypothesis.py:53-68) of ONE hypothesis (k views) against
// B candidate detections of camera `ocam`: cost [B], veto [B] ------------------------------------------
__global__ void k_hypothesis_cost(const float* __restrict__ F, int V, const double* __restrict__ hpose,
                                  const int* __restrict__ hcam, int k, const double* __restrict__ opose, int ocam, int B,
                                  int J, double epi_thr, double veto_believe, double* __restrict__ cost,
                                  unsigned char* __restrict__ veto) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* o = opose + (int64_t)b * J * 3;
    double kept[PAM_MAX_J];
    int nk = 0;
    for (int j = 0; j < J; ++j) if (o[j * 3 + 2] >= 0.0) kept[nk++] = o[j * 3 + 2];
    const double believe = np_sum(kept, nk) / (double)nk;
    double total = 0.0;
    bool vt = false;
    for (int q = 0; q < k; ++q) {
        double Fm[9];
        load9(F + ((int64_t)hcam[q] * V + ocam) * 9, Fm);
        const double* p = hpose + (int64_t)q * J * 3;
        NpSumStream<double> acc;
        acc.begin(J);
        for (int j = 0; j < J; ++j) {
            double d1, d2;
            epi_pair_cv(Fm, p[j * 3 + 1], p[j * 3], o[j * 3 + 1], o[j * 3], d1, d2);
            acc.push((d1 * p[j * 3 + 2] + d2 * o[j * 3 + 2]) / 2.0);
        }
        const double pc = acc.total() / (double)J / epi_thr;
        total += pc;
        if (pc > 1.0 && believe > veto_believe) vt = true;
    }
    cost[b] = total / (double)k;
    veto[b] = vt ? 1 : 0;
}

// ---- section 8f-1: PCP / MPJPE counters of Evaluate3DPose_PCP (evalmodel.py:120-206) ------------------
// One thread per (sequence, frame, ground-truth actor).  Predicted poses come straight from the
// tracker's output tensors (count [S][T], joints [S][T][MT][J][3] f32); gt [S][T][P][14][3] f64 in
// Shelf/Campus joint order, gt_valid [S][T][P].  remap 0: predictions already have the 14 Shelf joints;
// remap 1: COCO-17 -> Shelf-14 with coco2shelf3D (eval/transformation.py:5-39).
// counters [P][10][2] int64 = (correct, evaluated) per actor and bone (9 limbs + hip-head), summed over
// all sequences / frames with atomics; mpjpe [2] f64 = (sum of per-joint errors, joints counted).
__device__ __forceinline__ void to_shelf14(const float* __restrict__ p, int remap, double (*o)[3]) {
    if (!remap) {
        for (int j = 0; j < 14; ++j) for (int k = 0; k < 3; ++k) o[j][k] = (double)p[j * 3 + k];
        return;
    }
    const int map[12] = {16, 14, 12, 11, 13, 15, 10, 8, 6, 5, 7, 9};
    for (int j = 0; j < 12; ++j) for (int k = 0; k < 3; ++k) o[j][k] = (double)p[map[j] * 3 + k];
    const double top[3] = {0.78, 0.5, 1.5}, bot[3] = {0.3, 0.4, 0.6};
    for (int k = 0; k < 3; ++k) {
        const double mid = (o[8][k] + o[9][k]) / 2.0, nose = (double)p[k];
        o[13][k] = mid + (nose - mid) * top[k];
        o[12][k] = mid + (nose - mid) * bot[k];
    }
}
__device__ __forceinline__ double dist3(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt(x * x + y * y + z * z);
}
// The same with the short inline square root (MUFU seed, two Newton steps, residual correction: faithfully rounded,
// no call and no slow-path branch), so that the 26 independent distances of one pose interleave instead of running
// one after the other (k_eval_pcp is bound by exactly that latency).  A limb test would need `<=` to hold within one
// ulp to see the difference.
__device__ __forceinline__ double dist3_inline(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return sqrt_f64(x * x + y * y + z * z);
}
// joint j of to_shelf14's mapping, alone (same expressions, so the same bits)
__device__ __forceinline__ void shelf14_joint(const float* __restrict__ p, int remap, int j, double* o) {
    if (!remap) {
        for (int k = 0; k < 3; ++k) o[k] = (double)p[j * 3 + k];
        return;
    }
    // {16, 14, 12, 11, 13, 15, 10, 8, 6, 5, 7, 9} packed in nibbles - 5 (no local array, no constant bank)
    const int m = 5 + (int)((0xB9768A531024ull >> (4 * (11 - j))) & 15ull);
    if (j < 12) {
        for (int k = 0; k < 3; ++k) o[k] = (double)p[m * 3 + k];
        return;
    }
    const double top[3] = {0.78, 0.5, 1.5}, bot[3] = {0.3, 0.4, 0.6};
    for (int k = 0; k < 3; ++k) {
        const double mid = ((double)p[6 * 3 + k] + (double)p[5 * 3 + k]) / 2.0, nose = (double)p[k];
        o[k] = mid + (nose - mid) * (j == 13 ? top[k] : bot[k]);
    }
}
#define PAM_EVAL_MAX_P 64
#define PAM_EVAL_THREADS 128
#define PAM_EVAL_GT_STRIDE 42      // doubles per staged ground-truth pose: rows stay contiguous, as in HBM (one bulk copy)
// Dynamic shared memory: the block's 128 ground-truth poses [128][43] f64, then the predicted poses of the frames
// they belong to [frames][MT][J][3] f32.  Both are contiguous in HBM, so they are staged with fully coalesced loads
// (the kernel is bandwidth bound: 1344 B of ground truth + up to MT x J x 12 B of predictions per frame); the
// arithmetic then runs out of shared memory.
__host__ __device__ inline int eval_frames_per_block(int P) { return (PAM_EVAL_THREADS + P - 1) / P + 1; }
__host__ __device__ inline size_t eval_smem_bytes(int P, int MT, int J) {
    // (+ 32: the predictions are bulk-copied from the 16-byte boundary below their first byte, in 16-byte units)
    return (size_t)PAM_EVAL_THREADS * PAM_EVAL_GT_STRIDE * 8 + ((size_t)eval_frames_per_block(P) * MT * J * 3 * 4 + 15) / 16 * 16 + 32 +
           (size_t)P * 20 * 4;          // + the block's counters: 72 KB for the Shelf shape, three blocks per SM
}
__global__ void __launch_bounds__(PAM_EVAL_THREADS)
k_eval_pcp(const int* __restrict__ count, const float* __restrict__ joints, const double* __restrict__ gt,
           const unsigned char* __restrict__ gt_valid, int S, int T, int P, int MT, int J, int remap,
           int t0, int t1, double alpha, unsigned long long* __restrict__ counters, double* __restrict__ mpjpe) {
    extern __shared__ __align__(128) unsigned char eval_smem[];
    double* s_gt = (double*)eval_smem;
    unsigned char* s_pred_raw = (unsigned char*)(s_gt + PAM_EVAL_THREADS * PAM_EVAL_GT_STRIDE);      // 16-byte aligned
    // block-local counters first (shared-memory atomics), one global atomic per counter per block
    unsigned int* s_cnt = (unsigned int*)(eval_smem + eval_smem_bytes(P, MT, J) - (size_t)P * 20 * 4);
    __shared__ double s_err[4];
    __shared__ unsigned int s_nj;
    __shared__ __align__(8) unsigned long long s_bar;
    for (int i = threadIdx.x; i < P * 20; i += blockDim.x) s_cnt[i] = 0u;
    if (threadIdx.x < 4) s_err[threadIdx.x] = 0.0;
    if (threadIdx.x == 0) s_nj = 0u;
    const int64_t total = (int64_t)S * T * P;
    const int64_t item0 = (int64_t)blockIdx.x * PAM_EVAL_THREADS;
    const int nitems = (int)((total - item0) < PAM_EVAL_THREADS ? (total - item0) : PAM_EVAL_THREADS);
    const int64_t frame0 = item0 / P, frame1 = (item0 + nitems - 1) / P;       // flattened (sequence, frame) indices
    const int nframes = (int)(frame1 - frame0 + 1);
    // this thread's item: its scalar inputs are requested before the staging barrier, their latency overlaps the copies
    const int64_t it = item0 + threadIdx.x;
    const int local = (int)(item0 - frame0 * P) + (int)threadIdx.x;            // item index relative to the block's first frame
    const int df = local / P, pid = local - df * P;
    const int64_t st = frame0 + df;
    int t = (int)(frame0 % T) + df;
    while (t >= T) t -= T;
    const bool live = it < total && t >= t0 && t < t1;
    const bool valid = live && gt_valid[it];
    const int cnt_st = live ? count[st] : 0;
    const double* src = gt + item0 * 42;
    // predictions of the block's frames: nframes x MT x J x 3 contiguous floats, starting anywhere on a 4-byte boundary
    const int row = MT * J * 3;
    const float* ps = joints + frame0 * row;
    const int n = nframes * row;
    const unsigned mis = (unsigned)(((uintptr_t)ps) & 15);
    float* s_pred = (float*)(s_pred_raw + mis);
    // Bulk copies (TMA): the ground truth as it is (336 B per pose: every block offset is a multiple of 16), the
    // predictions from the 16-byte boundary below their first byte, rounded up to 16-byte units -- except in the last
    // block, where that could read past the end of the tensor.
    const bool bulk = (((uintptr_t)src) & 15) == 0 && blockIdx.x + 1 < gridDim.x;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_bar);
    if (bulk && threadIdx.x == 0) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(s_gt), bytes = (unsigned)nitems * 336u;
        const unsigned pdst = (unsigned)__cvta_generic_to_shared(s_pred_raw), pbytes = ((unsigned)n * 4u + mis + 15u) & ~15u;
        const unsigned char* psrc = (const unsigned char*)ps - mis;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes + pbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(pdst), "l"(psrc), "r"(pbytes), "r"(bar) : "memory");
    }
    if (!bulk) {
        for (int i = threadIdx.x; i < nitems * 42; i += PAM_EVAL_THREADS) s_gt[i] = src[i];
        for (int i = threadIdx.x; i < n; i += PAM_EVAL_THREADS) s_pred[i] = ps[i];
    }
    __syncthreads();
    if (bulk)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "PAM_EVAL_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@!p bra PAM_EVAL_WAIT_%=;\n\t}" ::"r"(bar) : "memory");
    double e = 0.0;
    bool scored = false;
    {
        if (valid) {
            unsigned int* cnt = s_cnt + pid * 20;
            const int k = cnt_st < MT ? cnt_st : MT;            // rows beyond the output stride were not written
            if (k <= 0) {                  // "Cannot get any pose in frame": all ten parts count as errors
                for (int b = 0; b < 10; ++b) atomicAdd(cnt + b * 2 + 1, 1u);
            } else {
                const double* g = s_gt + threadIdx.x * PAM_EVAL_GT_STRIDE;
                const float* pf = s_pred + (int64_t)(st - frame0) * MT * J * 3;
                // The mapped prediction is never kept whole (42 doubles per thread): joints are re-read from shared
                // memory where they are needed.  Four candidates are summed side by side (independent chains, each in
                // the reference's order), and every joint distance is taken once -- the limb tests and the MPJPE sum
                // use the same 14 values (evalmodel.py:176-200 computes them twice, to the same bits).
                double best = 0.0;
                int bq = 0;
                for (int q0 = 0; q0 < k; q0 += 4) {
                    double d[4] = {0.0, 0.0, 0.0, 0.0};      // vectorize_distance: squared distance over all 42 coordinates
                    const float* pq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) pq[u] = pf + (q0 + u < k ? q0 + u : k - 1) * J * 3;
#pragma unroll 2
                    for (int j = 0; j < 14; ++j) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            double m[3];
                            shelf14_joint(pq[u], remap, j, m);
                            for (int c = 0; c < 3; ++c) { const double x = g[j * 3 + c] - m[c]; d[u] += x * x; }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (q0 + u < k && ((q0 + u) == 0 || d[u] < best)) { best = d[u]; bq = q0 + u; }
                }
                const float* pb = pf + bq * J * 3;           // the closest prediction
                double dj[14], m2[3], m3[3];
#pragma unroll
                for (int j = 0; j < 14; ++j) {
                    double m[3];
                    shelf14_joint(pb, remap, j, m);
                    if (j == 2) { m2[0] = m[0]; m2[1] = m[1]; m2[2] = m[2]; }
                    if (j == 3) { m3[0] = m[0]; m3[1] = m[1]; m3[2] = m[2]; }
                    dj[j] = dist3_inline(g + j * 3, m);
                }
                const int bones[9][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {9, 10}, {10, 11}, {12, 13}};
#pragma unroll
                for (int b = 0; b < 9; ++b) {
                    const int s0 = bones[b][0], e0 = bones[b][1];
                    const double len = dist3_inline(g + e0 * 3, g + s0 * 3);
                    if ((dj[s0] + dj[e0]) / 2.0 <= alpha * len) atomicAdd(cnt + b * 2, 1u);
                    atomicAdd(cnt + b * 2 + 1, 1u);
                }
                double gh[3], mh[3];
                for (int c = 0; c < 3; ++c) { gh[c] = (g[2 * 3 + c] + g[3 * 3 + c]) / 2.0; mh[c] = (m2[c] + m3[c]) / 2.0; }
                const double len = dist3_inline(g + 12 * 3, gh);
                if ((dist3_inline(gh, mh) + dj[12]) / 2.0 <= alpha * len) atomicAdd(cnt + 18, 1u);
                atomicAdd(cnt + 19, 1u);
#pragma unroll
                for (int j = 0; j < 14; ++j) e += dj[j];
                scored = true;
            }
        }
    }
    // MPJPE: warp shuffle reduction, then one shared slot per warp
    unsigned int nsc = __ballot_sync(0xffffffffu, scored);
    for (int off = 16; off > 0; off >>= 1) e += __shfl_down_sync(0xffffffffu, e, off);
    if ((threadIdx.x & 31) == 0) { s_err[threadIdx.x >> 5] = e; atomicAdd(&s_nj, 14u * __popc(nsc)); }
    __syncthreads();
    for (int i = threadIdx.x; i < P * 20; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(counters + i, (unsigned long long)s_cnt[i]);
    if (threadIdx.x == 0 && s_nj) {
        atomicAdd(mpjpe, s_err[0] + s_err[1] + s_err[2] + s_err[3]);
        atomicAdd(mpjpe + 1, (double)s_nj);
    }
}

// ---- section 8f-1: EvaluatePanoptic.evaluate matching (evalmodel.py:291-320) -----------------------------
// One thread per (sequence, frame, predicted pose): the prediction (tracker output, metres) is mapped
// to the 14 evaluated Panoptic joints in millimetres -- COCO-17: nose, mid-hip = (l-hip + r-hip)/2, then
// [5,7,9,11,13,15,6,8,10,12,14,16]; COCO-19: joints 1..14 -- and compared with every ground-truth body
// of the frame: MPJPE over the visible joints, minimum and arg-minimum over the bodies.
// gt [S][T][G][14][3] f64 (mm), vis [S][T][G][14] u8, n_gt [S][T];  out_mpjpe / out_gt [S][T][MT]
// (out_gt = -1: frame without ground truth or slot beyond count).
__global__ void k_eval_panoptic_match(const int* __restrict__ count, const float* __restrict__ joints,
                                      const double* __restrict__ gt, const unsigned char* __restrict__ vis,
                                      const int* __restrict__ n_gt, int S, int T, int G, int MT, int J,
                                      double* __restrict__ out_mpjpe, int* __restrict__ out_gt) {
    const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= (int64_t)S * T * MT) return;
    const int q = (int)(it % MT);
    const int64_t st = it / MT;
    out_gt[it] = -1;
    out_mpjpe[it] = 0.0;
    const int ng = n_gt[st];
    if (q >= count[st] || ng <= 0) return;
    const float* p = joints + it * J * 3;
    double m[14][3];
    if (J == 17) {
        const int map[12] = {5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16};
        for (int c = 0; c < 3; ++c) {
            m[0][c] = (double)p[c] * 1000.0;
            m[1][c] = ((double)p[11 * 3 + c] * 1000.0 + (double)p[12 * 3 + c] * 1000.0) / 2.0;
        }
        for (int j = 0; j < 12; ++j) for (int c = 0; c < 3; ++c) m[2 + j][c] = (double)p[map[j] * 3 + c] * 1000.0;
    } else {
        for (int j = 0; j < 14; ++j) for (int c = 0; c < 3; ++c) m[j][c] = (double)p[(j + 1) * 3 + c] * 1000.0;
    }
    double best = 0.0;
    int arg = -1;
    for (int g = 0; g < ng; ++g) {
        const double* gg = gt + ((int64_t)st * G + g) * 14 * 3;
        const unsigned char* vv = vis + ((int64_t)st * G + g) * 14;
        double d[14];
        int nvz = 0;
        for (int j = 0; j < 14; ++j)
            if (vv[j]) {
                const double x = m[j][0] - gg[j * 3], y = m[j][1] - gg[j * 3 + 1], z = m[j][2] - gg[j * 3 + 2];
                d[nvz++] = sqrt(x * x + y * y + z * z);
            }
        const double mp = np_sum(d, nvz) / (double)nvz;
        if (arg < 0 || mp < best) { best = mp; arg = g; }      // np.argmin: first minimum; NaN never wins
    }
    out_mpjpe[it] = best;
    out_gt[it] = arg;
}

// ---- section 8f-4: the reference's alternative smoother and pair-wise triangulation (no live caller) ----------

// One-Euro filter bank (tracking/OneEuroFilter.py:12-77): n channels that share their time stamps (the way IterTrack
// holds one filter per joint coordinate, tracking/IterativeTracker.py:231-237).  state [n][4] f64 = {last raw value,
// filtered value, filtered derivative, initialised}; `freq` is maintained by the host exactly like the reference
// (1 / (t - t_last) once both are truthy).  IEEE operations in the reference's order, no FMA contraction: bit-identical
// to the python class.
__device__ __forceinline__ double one_euro_alpha(double freq, double cutoff) {
    const double te = __ddiv_rn(1.0, freq);
    const double tau = __ddiv_rn(1.0, __dmul_rn(6.283185307179586, cutoff));      // 2 * math.pi * cutoff
    return __ddiv_rn(1.0, __dadd_rn(1.0, __ddiv_rn(tau, te)));
}
__global__ void k_one_euro(int n, const double* __restrict__ x, double freq, double mincutoff, double beta, double dcutoff,
                           double* __restrict__ state, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* st = state + (int64_t)i * 4;
    const bool have = st[3] != 0.0;
    const double xi = x[i];
    const double dx = have ? __dmul_rn(__dsub_rn(xi, st[0]), freq) : 0.0;
    const double ad = one_euro_alpha(freq, dcutoff);
    const double edx = have ? __dadd_rn(__dmul_rn(ad, dx), __dmul_rn(__dsub_rn(1.0, ad), st[2])) : dx;
    const double cutoff = __dadd_rn(mincutoff, __dmul_rn(beta, fabs(edx)));
    const double a = one_euro_alpha(freq, cutoff);
    const double s = have ? __dadd_rn(__dmul_rn(a, xi), __dmul_rn(__dsub_rn(1.0, a), st[1])) : xi;
    st[0] = xi; st[1] = s; st[2] = edx; st[3] = 1.0;
    out[i] = s;
}

// top_down_pose_kernel (utils/construction.py:9-31): every camera pair triangulates all joints from its two views
// (the 4 x 4 homogeneous system of cv2.triangulatePoints: rows x P2 - P0, y P2 - P1 of both views, not normalised);
// the pair whose 3-D pose reprojects best into ALL cameras -- sum over cameras of the Frobenius norm of the (J, 2)
// residual, with the reference's "+ 10e-6" on the homogeneous depth -- wins.
// One block per batch item; poses2d [batch][V][J][2] (x, y), cam [batch][V]; out pose3d [batch][J][3], pair [batch][2].
__device__ __forceinline__ bool pair_triangulate(const float* __restrict__ P, int ca, int cb, const double* pa, const double* pb,
                                                 double* xh) {
    double Pa[12], Pb[12];
    load12(P + ca * 12, Pa);
    load12(P + cb * 12, Pb);
    DltAccum acc;
    acc.reset(true);
    acc.add_row(pa[0] * Pa[8] - Pa[0], pa[0] * Pa[9] - Pa[1], pa[0] * Pa[10] - Pa[2], pa[0] * Pa[11] - Pa[3]);
    acc.add_row(pa[1] * Pa[8] - Pa[4], pa[1] * Pa[9] - Pa[5], pa[1] * Pa[10] - Pa[6], pa[1] * Pa[11] - Pa[7]);
    acc.add_row(pb[0] * Pb[8] - Pb[0], pb[0] * Pb[9] - Pb[1], pb[0] * Pb[10] - Pb[2], pb[0] * Pb[11] - Pb[3]);
    acc.add_row(pb[1] * Pb[8] - Pb[4], pb[1] * Pb[9] - Pb[5], pb[1] * Pb[10] - Pb[6], pb[1] * Pb[11] - Pb[7]);
    return acc.solve_homog(xh);
}
__global__ void __launch_bounds__(128)
k_top_down(const float* __restrict__ P, const double* __restrict__ poses2d, const int* __restrict__ cam, int V, int J,
           double* __restrict__ out_pose, int* __restrict__ out_pair, double* __restrict__ out_err) {
    extern __shared__ double s_err[];                      // [pairs][V] squared residuals per (pair, camera)
    const int b = blockIdx.x;
    const double* p2 = poses2d + (int64_t)b * V * J * 2;
    const int* cm = cam + (int64_t)b * V;
    const int npairs = V * (V - 1) / 2;
    // (pair, camera): squared Frobenius residual of the pair's pose in that camera
    for (int it = threadIdx.x; it < npairs * V; it += blockDim.x) {
        const int pr = it / V, k = it - pr * V;
        int i = 0, rem = pr;
        while (rem >= V - 1 - i) { rem -= V - 1 - i; ++i; }
        const int j2 = i + 1 + rem;
        double Pk[12];
        load12(P + cm[k] * 12, Pk);
        double acc = 0.0;
        for (int j = 0; j < J; ++j) {
            double xh[4];
            if (!pair_triangulate(P, cm[i], cm[j2], p2 + (i * J + j) * 2, p2 + (j2 * J + j) * 2, xh)) { acc = HUGE_VAL; break; }
            const double a = Pk[0] * xh[0] + Pk[1] * xh[1] + Pk[2] * xh[2] + Pk[3] * xh[3];
            const double c2 = Pk[4] * xh[0] + Pk[5] * xh[1] + Pk[6] * xh[2] + Pk[7] * xh[3];
            const double w = Pk[8] * xh[0] + Pk[9] * xh[1] + Pk[10] * xh[2] + Pk[11] * xh[3] + 10e-6;
            const double dx = a / w - p2[(k * J + j) * 2], dy = c2 / w - p2[(k * J + j) * 2 + 1];
            acc += dx * dx + dy * dy;
        }
        s_err[it] = acc;
    }
    __syncthreads();
    __shared__ int s_best;
    if (threadIdx.x == 0) {
        double best = 0.0;
        int arg = 0;
        for (int pr = 0; pr < npairs; ++pr) {
            double e = 0.0;
            for (int k = 0; k < V; ++k) e += sqrt(s_err[pr * V + k]);
            if (out_err) out_err[(int64_t)b * npairs + pr] = e;
            if (pr == 0 || e < best) { best = e; arg = pr; }       // np.argmin: first minimum
        }
        s_best = arg;
    }
    __syncthreads();
    int i = 0, rem = s_best;
    while (rem >= V - 1 - i) { rem -= V - 1 - i; ++i; }
    const int j2 = i + 1 + rem;
    if (threadIdx.x == 0 && out_pair) { out_pair[b * 2] = i; out_pair[b * 2 + 1] = j2; }
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        double xh[4];
        double* o = out_pose + ((int64_t)b * J + j) * 3;
        if (pair_triangulate(P, cm[i], cm[j2], p2 + (i * J + j) * 2, p2 + (j2 * J + j) * 2, xh)) {
            o[0] = xh[0] / xh[3]; o[1] = xh[1] / xh[3]; o[2] = xh[2] / xh[3];
        } else {
            o[0] = o[1] = o[2] = nan("");
        }
    }
}

// Kalman smoother bank (tracking/KalmanFilter.py:4-65: cv2.KalmanFilter(9, 3), constant-acceleration model per 3-D
// joint, float32).  One thread per filter; state [n][90] f32 = statePre/Post (9) + error covariance (81).
// step = the reference's predict(pt3d): correct(measurement) when one is given, then predict(); returns the predicted
// position.  Products are accumulated in double and rounded to float32 per matrix like OpenCV's float gemm; the 3 x 3
// innovation covariance is inverted directly (OpenCV: SVD solve), so results agree to float32 rounding, not bit for bit.
struct Kalman9 {
    float v, a, q, r;      // dt, dt^2 / 2, process noise (0.007), measurement noise (0.1)
};
__device__ __forceinline__ float kal_A(const Kalman9& k, int i, int j) {      // transition matrix, row i, column j
    if (i == j) return 1.f;
    if (j == i + 3) return k.v;
    if (j == i + 6) return k.a;
    return 0.f;
}
__global__ void k_kalman9(int n, Kalman9 kp, const double* __restrict__ meas /* [n][3] or null */,
                          const unsigned char* __restrict__ has_meas /* [n] or null = all */, float* __restrict__ state,
                          double* __restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    float* x = state + (int64_t)f * 90;
    float* P = x + 9;
    if (meas && (!has_meas || has_meas[f])) {
        // correct(): H = first three rows of A (tracking/KalmanFilter.py:13-17)
        float T2[3][9], S[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0.0;
                for (int k = 0; k < 9; ++k) s += (double)kal_A(kp, i, k) * (double)P[k * 9 + j];
                T2[i][j] = (float)s;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
                for (int k = 0; k < 9; ++k) s += (double)T2[i][k] * (double)kal_A(kp, j, k);
                S[i][j] = (float)(s + (i == j ? (double)kp.r : 0.0));
            }
        // temp4 = S^-1 T2  (3 x 9);  gain = temp4^T
        double Sd[9], Si[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Sd[i * 3 + j] = (double)S[i][j];
        inv33<double>(Sd, Si);
        float G[9][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += Si[i * 3 + k] * (double)T2[k][j];
                G[j][i] = (float)s;
            }
        float y[3];
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int k = 0; k < 9; ++k) s += (double)kal_A(kp, i, k) * (double)x[k];
            y[i] = (float)((double)(float)meas[(int64_t)f * 3 + i] - s);
        }
        float xn[9], Pn[81];
        for (int i = 0; i < 9; ++i) {
            double s = (double)x[i];
            for (int k = 0; k < 3; ++k) s += (double)G[i][k] * (double)y[k];
            xn[i] = (float)s;
        }
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j) {
                double s = (double)P[i * 9 + j];
                for (int k = 0; k < 3; ++k) s -= (double)G[i][k] * (double)T2[k][j];
                Pn[i * 9 + j] = (float)s;
            }
        for (int i = 0; i < 9; ++i) x[i] = xn[i];
        for (int i = 0; i < 81; ++i) P[i] = Pn[i];
    }
    // predict(): x = A x;  P = A P A^T + Q
    float xp[9], T1[81];
    for (int i = 0; i < 9; ++i) {
        double s = 0.0;
        for (int k = 0; k < 9; ++k) s += (double)kal_A(kp, i, k) * (double)x[k];
        xp[i] = (float)s;
    }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) {
            double s = 0.0;
            for (int k = 0; k < 9; ++k) s += (double)kal_A(kp, i, k) * (double)P[k * 9 + j];
            T1[i * 9 + j] = (float)s;
        }
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) {
            double s = 0.0;
            for (int k = 0; k < 9; ++k) s += (double)T1[i * 9 + k] * (double)kal_A(kp, j, k);
            P[i * 9 + j] = (float)(s + (i == j ? (double)kp.q : 0.0));
        }
    for (int i = 0; i < 9; ++i) x[i] = xp[i];
    out[(int64_t)f * 3] = (double)xp[0]; out[(int64_t)f * 3 + 1] = (double)xp[1]; out[(int64_t)f * 3 + 2] = (double)xp[2];
}

}  // namespace pam
