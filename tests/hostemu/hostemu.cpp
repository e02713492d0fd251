// TEST / DEBUG INFRASTRUCTURE ONLY.
// Compiles the block-cooperative tracker source (csrc/pam_track.h) for the HOST with a single
// "thread" so that the algorithmic logic of the CUDA kernel can be checked against the oracle in
// the GPU-less build container.  It is built by tests/hostemu/build.py into tests/hostemu/_build/,
// is loaded only by tests/ (-m "not gpu"), and is never imported, linked or shipped by the
// package: the product path is libpam.so and fails loudly without a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc/pam_host.h"

using namespace pam;
bool pam::HostCtx::kTwoPassAffinity = false;

template <class K>
static int run_sequences(const DevCfg& c, const CamConst& cc, char* state, int S, int T, int frame0, const float* dets,
                         const int32_t* counts, int32_t* out_count, int32_t* out_ids, float* out_joints,
                         uint8_t* out_nviews, int32_t* out_assoc, int32_t* status, uint8_t* out_vlist) {
    std::vector<double> arena((size_t)c.arena_bytes / 8 + 16);
    std::vector<CamShared<K>> cam(1);
    HostCtx ctx;
    NoHook hook;
    load_cameras(0, 1, c, cam.data(), cc);
    const int64_t fstride = (int64_t)c.V * c.D * c.J * 3;
    for (int s = 0; s < S; ++s) {
        std::memset((void*)arena.data(), 0, arena.size() * 8);
        Seq<K> sq;
        sq.bind(c, (char*)arena.data(), cam.data(), state + (int64_t)s * c.seq_bytes);
        load_state(ctx, c, sq);
        for (int t = 0; t < T; ++t) {
            const int64_t ft = (int64_t)s * T + t;
            FrameOut o;
            o.count = out_count + ft;
            o.ids = out_ids ? out_ids + ft * c.max_rep : nullptr;
            o.joints = out_joints ? out_joints + ft * c.max_rep * c.J * 3 : nullptr;
            o.nviews = out_nviews ? out_nviews + ft * c.max_rep * c.J : nullptr;
            o.assoc = out_assoc ? out_assoc + ft * c.V * c.D : nullptr;
            o.timing = nullptr;
            o.vlist = out_vlist ? out_vlist + ft * c.max_rep * PAM_VLIST : nullptr;
            frame_step(ctx, c, sq, frame0 + t, dets + ft * fstride, counts + ft * c.V, o,
                       dets + (int64_t)s * T * fstride, frame0, hook);
        }
        persist_views(ctx, c, sq, dets + (int64_t)s * T * fstride, frame0, (const float*)nullptr, -1);
        store_state(ctx, c, sq);
        if (status) status[s] = sq.sh->hdr.status | (sq.sh->hdr.warn << 8);
    }
    return PAM_OK;
}

extern "C" int hostemu_track_sequences(const pam_config* cfg, const float* P, const float* RKinv, const double* pos,
                                       const float* F, int S, int T, int frame0, const float* dets,
                                       const int32_t* counts, int32_t* out_count, int32_t* out_ids,
                                       float* out_joints, uint8_t* out_nviews, int32_t* out_assoc,
                                       int32_t* status, void* state_io /* may be null; S*seq_bytes, zero = fresh */,
                                       uint8_t* out_vlist /* may be null */) {
    DevCfg c;
    std::string err;
    // PAM_HOSTEMU_LEAN=1: the throughput flavour of the working set (one detection buffer, raw pose in the
    // sequence's global scratch); default: the latency flavour.  PAM_HOSTEMU_CAPS=0/1/2 forces a capacity class
    // at least that large (the same configuration must give the same result in every class that holds it).
    const char* lean = getenv("PAM_HOSTEMU_LEAN");
    const bool is_lean = lean && lean[0] == '1';
    int rc = make_devcfg(*cfg, c, err, is_lean ? 1 : 2, !is_lean);
    if (rc != PAM_OK) return rc;
    if (!tracker_capable(*cfg)) return PAM_E_INVALID;
    const char* tp = getenv("PAM_HOSTEMU_TWOPASS");      // the throughput launches' two-pass affinity
    HostCtx::kTwoPassAffinity = tp && tp[0] == '1';
    const char* fc = getenv("PAM_HOSTEMU_CAPS");
    if (fc && atoi(fc) > c.caps) force_caps(c, atoi(fc));
    std::vector<char> own;
    char* state = (char*)state_io;
    if (!state) { own.assign((size_t)c.seq_bytes * S, 0); state = own.data(); }
    CamConst cc{P, RKinv, pos, F};
    if (c.caps == CAPS_SMALL)
        return run_sequences<CapsSmall>(c, cc, state, S, T, frame0, dets, counts, out_count, out_ids, out_joints, out_nviews, out_assoc, status, out_vlist);
    if (c.caps == CAPS_MID)
        return run_sequences<CapsMid>(c, cc, state, S, T, frame0, dets, counts, out_count, out_ids, out_joints, out_nviews, out_assoc, status, out_vlist);
    return run_sequences<CapsMax>(c, cc, state, S, T, frame0, dets, counts, out_count, out_ids, out_joints, out_nviews, out_assoc, status, out_vlist);
}

// decision margins of the sequences in `state` (PAM_MARGIN builds of the harness)
extern "C" int hostemu_margins(const pam_config* cfg, const void* state, int S, double* out) {
    DevCfg c;
    std::string err;
    int rc = make_devcfg(*cfg, c, err);
    if (rc != PAM_OK) return rc;
    for (int s = 0; s < S; ++s)
        std::memcpy(out + (size_t)s * MG_COUNT, (const char*)state + (int64_t)s * c.seq_bytes + c.off_margin, 8 * MG_COUNT);
    return PAM_OK;
}

extern "C" int hostemu_state_layout(const pam_config* cfg, pam_state_layout* L) {
    DevCfg c;
    std::string err;
    int rc = make_devcfg(*cfg, c, err);
    if (rc != PAM_OK) return rc;
    fill_layout(c, *L);
    return PAM_OK;
}

// the building blocks below phase level, one at a time (tests/test_primitives_hostemu.py)
extern "C" int hostemu_lsap(int nr, int nc, const double* cost /*[nr][nc]*/, int* col4row /*[nr]*/) {
    return pam::lsap_solve<64>(nr, nc, [&](int i, int j) { return cost[i * nc + j]; }, col4row);
}
extern "C" int hostemu_gaussian_weights(double sigma, double* w /*[PAM_MAX_RADIUS + 1]*/) { return pam::gaussian_weights(sigma, w); }
extern "C" int hostemu_reflect_index(int i, int n) { return pam::reflect_index(i, n); }
extern "C" double hostemu_np_sum(const double* x, int n) { return pam::np_sum(x, n); }
extern "C" double hostemu_mean_confidence(const float* pose /*[J][3]*/, int J) { return pam::mean_confidence(pose, J); }
extern "C" double hostemu_sqrt(double x) { return pam::sqrt_f64(x); }
extern "C" double hostemu_rsqrt(double x) { return pam::rsqrt_f64(x); }
extern "C" double hostemu_rcp(double x) { return pam::rcp_f64(x); }

// One camera's assignment problem through the kernel's own solvers (pam_track.h: quick test as in phase 3 of
// frame_step, then assign_closed_form, assign_enumerated, and the general solver when those decline).
// A [n][mm] affinities (n <= 32 tracks, mm <= 16 detections); t2d [n] = detection of each track or -1;
// returns which stage decided: 0 quick test, 1 closed form, 2 enumeration, 3 general solver.
extern "C" int hostemu_assign(const double* A, int n, int mm, int limit, int* t2d) {
    typedef pam::CapsMax K;
    static pam::SeqShared<K> sh;
    memset(&sh, 0, sizeof sh);
    memset(sh.match, -1, sizeof sh.match);
    for (int i = 0; i < n; ++i) for (int d = 0; d < mm; ++d) sh.aff[0][i][d] = A[i * mm + d];
    sh.m[0] = (signed char)mm;
    bool conflict = false;
    for (int i = 0; i < n; ++i) {                       // the quick test of phase 3
        int cnt = 0, arg = -1;
        for (int d = 0; d < mm; ++d) if (sh.aff[0][i][d] > 0.0) { ++cnt; arg = d; }
        if (cnt == 1) {
            int col = 0;
            for (int k = 0; k < n; ++k) col += (sh.aff[0][k][arg] > 0.0) ? 1 : 0;
            if (col == 1) { sh.t2d(0, i) = (signed char)arg; sh.d2t(0, arg) = (signed char)i; }
            else conflict = true;
        } else if (cnt > 1) conflict = true;
    }
    int how = 0;
    if (conflict) {
        pam::HostCtx ctx;
        if (pam::assign_closed_form(sh, 0, n, mm)) how = 1;
        else if (pam::assign_enumerated(ctx, sh, 0, n, mm, limit)) how = 2;
        else {
            how = 3;
            for (int i = 0; i < n; ++i) sh.t2d(0, i) = -1;
            int col4row[PAM_LSAP_N];
            pam::lsap_solve<PAM_LSAP_N>(n, mm, [&](int i, int d) { return -sh.aff[0][i][d]; }, col4row);
            for (int i = 0; i < n; ++i) if (col4row[i] >= 0 && sh.aff[0][i][col4row[i]] > 0.0) sh.t2d(0, i) = (signed char)col4row[i];
        }
    }
    for (int i = 0; i < n; ++i) t2d[i] = sh.t2d(0, i);
    return how;
}

// DLT extractor check: n joints, V views each; mode 0 = product policy (Gram when all weights are 1),
// 1 = force Givens + inverse iteration/Jacobi, 2 = force Givens + Jacobi only.
extern "C" int hostemu_dlt(int n, int V, const double* P, const double* uv, const double* w, const uint8_t* keep,
                           int mode, double* X, int* path) {
    for (int i = 0; i < n; ++i) {
        bool fresh = true;
        for (int a = 0; a < V; ++a) if (keep[i * V + a] && w[i * V + a] != 1.0) fresh = false;
        DltAccum acc;
        int how = -1;
        if (mode == 0) {
            acc.reset(true);
            for (int a = 0; a < V; ++a) if (keep[i * V + a]) acc.add_view(P + a * 12, uv[(i * V + a) * 2], uv[(i * V + a) * 2 + 1], w[i * V + a]);
            acc.solve(X + i * 3, &how);
            if (how >= 0) how += 10;
        }
        if (how < 0) {
            acc.reset(false);
            for (int a = 0; a < V; ++a) if (keep[i * V + a]) acc.add_view(P + a * 12, uv[(i * V + a) * 2], uv[(i * V + a) * 2 + 1], w[i * V + a]);
            if (mode == 2) { double x[4]; acc.jacobi(x); X[i*3] = x[0]/x[3]; X[i*3+1] = x[1]/x[3]; X[i*3+2] = x[2]/x[3]; how = 1; }
            else acc.solve(X + i * 3, &how);
        }
        path[i] = how;
    }
    return 0;
}

#if defined(PAM_COUNT_ITERS)
extern "C" void hostemu_invit_stats(long long* steps, long long* calls) { *steps = pam::g_invit_steps; *calls = pam::g_invit_calls; }
#endif
