// pam_track.h -- one frame of the part-aware tracker for ONE sequence, written group-cooperatively.
//
// frame_step<Ctx>() is executed by all threads of one thread GROUP (Ctx: 1..8 warps of a CTA; several
// groups = several sequences share a CTA and its camera constants) which owns one sequence; phases are
// separated by ctx.phase_sync() / ctx.sync() (a warp barrier for one-warp groups, a named barrier otherwise;
// phase_sync may span the CTA's sequences, see `convoy`).  Work inside
// a phase is a group-stride loop over independent items (camera x track x detection, track x joint,
// hypothesis x detection ...), the few inherently serial steps (list bookkeeping, assignment problems)
// run on one thread per problem.  With Ctx = HostCtx (one "thread", sync = no-op) the same source runs
// on the CPU under tests/hostemu/ for debugging only.
//
// The working set of a sequence is ONE byte arena (shared memory on the device) whose layout is computed
// from the configuration (cameras, detections, joints, track slots), so the capacity limits are run-time
// values and small configurations pay a small footprint (Shelf shape: ~8 KB, which lets 24-28
// sequences share an SM).
//
// Reference semantics (paths under /root/reference/src):
//   phase 1-4  tracking/IterativeTracker.py:124-167   ageing, reprojection affinity, LSAP, add_pose
//   phase 5-7  tracking/IterativeTracker.py:253-274, 305-395   per-track update, life-cycle
//   phase 8    tracking/IterativeTracker.py:52-113 + tracking/hypothesis.py:11-77   new-track init
//   output     tracking/IterativeTracker.py:178 + ivclabpose.py:265-287   reap + output contract
#pragma once
#include "pam_core.h"

// views / view pairs in flight per thread in the part-aware filter and the Gram fold (phase 5)
#if !defined(PAM_P5_UNROLL)
#define PAM_P5_UNROLL 1
#endif

namespace pam {

enum { ST_TENTATIVE = 1, ST_CONFIRMED = 2, ST_DELETED = 3 };
// hard errors (sticky, the sequence stops reporting): only conditions make_devcfg() already excludes
enum { SEQ_OK = 0, SEQ_ERR_HIST_OVERFLOW = 4 };
// capacity warnings (NOT sticky: the excess is dropped for that frame, tracking goes on; the reference
// has no such limits, so a run that raised one may differ from it from that frame on)
enum {
    WARN_TRACK_OVERFLOW = 1,   // a new track found no free slot (cfg.max_trk) and was not created
    WARN_HYP_OVERFLOW = 2,     // more hypotheses than cfg.max_hyp during new-track initialisation
    WARN_DET_OVERFLOW = 4,     // counts[c] > cfg.D: the detections beyond D were ignored
};
// decision margins recorded by -DPAM_MARGIN builds: the smallest distance any decision of the run came to
// flipping (index into the per-sequence margin block)
enum {
    MG_ASSOC_C = 0,      // |c| of "c > 0" per (track, detection, joint)            IterativeTracker.py:143-146
    MG_JOINT_A = 1,      // |A| of "A < 0" per view pair and joint (update mode)     utils/matching.py:248
    MG_RAY = 2,          // |ra - rb| / max(ra, rb) of the conflict resolution        utils/matching.py:272-277
    MG_BELIEVE = 3,      // |believe - conf_threshold|                                IterativeTracker.py:59
    MG_INIT_A = 4,       // |A| (float32) of "A < 0" in init mode                     utils/matching.py:287
    MG_INIT_ROWSUM = 5,  // |s1 - s2| of the row-sum rule                             utils/matching.py:289-294
    MG_VETO = 6,         // |pose_cost - 1| where believe > veto threshold            tracking/hypothesis.py:66
    MG_ASSIGN = 7,       // smallest positive affinity entering an assignment
    MG_COUNT = 8
};

// Launch-constant parameters (kernel argument; lives in the constant bank).
struct DevCfg {
    int V, J, D, max_trk, max_hyp;
    int max_rep;                             // rows per frame of the output tensors (<= max_trk)
    float inv_J, inv_D, inv_V, inv_VD;       // reciprocals for fast_div
    int n_init, max_age, min_valid, stale_window;
    uint32_t arm_mask;
    int rad[2];                              // [0] sigma, [1] arm_sigma
    double gw[2][PAM_MAX_RADIUS + 1];
    double w_age[PAM_MAX_AGEW];              // exp(-lambda_t * T), T = 0..stale_window
    double inv_joint_thr, joint_thr2;
    double inv_denom_tab[16];                // 1 / (alpha2d * dt), dt < 16
    double inv_decay_tab[16];                // 1 / exp(lambda_a * dt)
    double conf_thr, epi_thr, joint_thr, alpha2d, lambda_a, veto_believe, fail_limit;
    float init_thr_f32;
    // per-sequence global state (bytes) -- see state_layout()
    int64_t off_hdr, off_meta, off_view, off_hist, off_vel, off_nv, off_init, off_raw, off_margin, seq_bytes;
    // per-sequence working arena: [SeqShared<K> (compile-time capacity class K)][detection buffers][raw pose]
    int caps;                                // capacity class (CAPS_*), chosen from V, D, J, max_trk
    int convoy;                              // one-warp groups of a CTA start every frame together (instruction-cache locality)
    int aff_probe;                           // two-pass affinity: joints evaluated before hopeless pairs are dropped (0 = off)
    int nbuf;                                // detection buffers: 2 = next frame staged during the current one
    int frame_floats;                        // V * D * J * 3, padded to a multiple of 4
    int a_dets, a_raw;                       // byte offsets of the run-time sized tail (a_raw < 0: raw pose in HBM scratch)
    int arena_bytes;
};

struct SeqHeader {
    int ntracks, next_id, status, frames_done;
    uint32_t used_mask;
    int warn;                                // WARN_* bits seen so far
    int warn_frames;                         // frames on which something was dropped
    int reserved;
    signed char order[PAM_MAX_TRK];          // track slots in track-list order
};

struct TrkMeta {
    int track_id, hits, age, tsu, state, already, nviews, hist_start, hist_len;
    int vt_last;                        // usable views of the last successful update (length of its joints_views)
    int view_cid[PAM_MAX_V];
    int view_time[PAM_MAX_V];
    int hist_time[PAM_HIST];
    signed char view_slot[PAM_MAX_V];   // camera -> position in the view list (dict key lookup), -1 = absent
    signed char view_det[PAM_MAX_V];    // detection index of the view inside its own frame (lazy persistence)
};

// scalars of the frame in flight
struct FrameScalars {
    int n;                 // tracks alive at frame start
    int any_conflict;      // some camera needs the full assignment solver
    int any_deleted;       // a track was deleted this frame: the track list needs compaction
    int do_init;
    int hyp_n;
    int warned;            // something was dropped this frame
    int n_surv;            // two-pass affinity: (track, camera, detection) items that survived the probe
    int unsure;            // sign-only affinity: some joint fell inside the guard band, the frame is evaluated exactly
    int bf_k;              // enumerated assignment: index of the best candidate
    unsigned long long bf_best;   //                  and the bit pattern of its total affinity
    unsigned long long bf_sig_lo, bf_sig_hi;   //     smallest / largest signature among the candidates that reach it
};

// One usable view of a track for the current frame: where its (v, u, conf) triples live (the staged
// detections for a view matched this frame, the launch's input or the persisted copy in HBM for a stale
// one), its camera and age.
struct ViewSrc {
    const float* p;
    int cid, T;
};

inline void state_layout(DevCfg& c) {
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t at = o; o += (bytes + 15) / 16 * 16; return at; };
    c.off_hdr = take(sizeof(SeqHeader));
    c.off_meta = take((int64_t)sizeof(TrkMeta) * c.max_trk);
    c.off_hist = take((int64_t)8 * c.max_trk * PAM_HIST * c.J * 3);
    c.off_view = take((int64_t)4 * c.max_trk * c.V * c.J * 3);
    c.off_vel = take((int64_t)4 * c.max_trk * c.J * 3);
    c.off_nv = take((int64_t)c.max_trk * c.J);
    // scratch of the (rare) new-track initialisation: hyp_pose, hyp_cost (f64), hyp_veto, hyp_nvj (u8)
    c.off_init = take((int64_t)8 * c.max_hyp * (c.J * 3 + c.D) + (int64_t)c.max_hyp * (c.D + c.J));
    // raw (unsmoothed) pose of the frame in flight when the launch shape keeps it out of shared memory
    c.off_raw = take((int64_t)8 * c.max_trk * c.J * 3);
    c.off_margin = take((int64_t)8 * MG_COUNT);
    c.seq_bytes = (o + 127) / 128 * 128;
}

// Capacity classes: the STRIDES of every array of the working set are compile-time constants (so indexing is
// shifts / immediates and everything is addressed as shared memory), the COUNTS (cameras, detections, joints,
// track slots) stay run-time values below them.  A configuration runs in the smallest class that holds it.
template <int CV, int CD, int CJ, int CT>
struct Caps {
    enum { V = CV, D = CD, J = CJ, T = CT, H = (CV * CD < PAM_MAX_HYP ? CV * CD : PAM_MAX_HYP) };
};
typedef Caps<5, 4, 17, 8> CapsSmall;      // Campus / Shelf shaped (also the reference-native 17 joints)
typedef Caps<5, 8, 19, 12> CapsMid;       // Panoptic shaped
typedef Caps<PAM_MAX_V, PAM_MAX_D, PAM_MAX_J, PAM_MAX_TRK> CapsMax;   // anything the tracker accepts
enum { CAPS_SMALL = 0, CAPS_MID = 1, CAPS_MAX = 2 };
inline int caps_class(int V, int D, int J, int T) {
    if (V <= CapsSmall::V && D <= CapsSmall::D && J <= CapsSmall::J && T <= CapsSmall::T) return CAPS_SMALL;
    if (V <= CapsMid::V && D <= CapsMid::D && J <= CapsMid::J && T <= CapsMid::T) return CAPS_MID;
    return CAPS_MAX;
}

// camera constants of the rig, widened to double, shared by every sequence of a CTA
// (every matrix starts on a 16-byte boundary, so that its rows are fetched with 128-bit shared-memory loads)
template <class K>
struct alignas(16) CamShared {
    double P[K::V][12];
    double RK[K::V][10];
    double pos[K::V][4];
    double F[K::V][K::V][10];
};

// The fixed part of one sequence's working set (shared memory on the device).
template <class K>
struct SeqShared {
    unsigned long long mbar[2];              // transaction barriers of the detection buffers
    int cnt[2][PAM_MAX_V];                   // staged per-camera detection counts
    SeqHeader hdr;
    FrameScalars fs;
    TrkMeta trk[K::T];
    double aff[K::V][K::T][K::D];
    double believe[K::V][K::D];              // mean confidence of every detection
    double inv_denom[K::T];                  // 1 / (alpha2d * dt)
    double inv_decay[K::T];                  // 1 / exp(lambda_a * dt)
    double thr2_lo[K::T], thr2_hi[K::T];     // (alpha2d dt)^2 (1 -+ 2^-24): a joint is surely inside / surely outside
    ViewSrc vsrc[K::T][K::V];                // gathered views per track, dict order
    int dt[K::T];
    int fail[K::T];                          // joints left with < 2 views
    // t2d [V][T] then d2t [V][D]: reset together, as ints
    signed char match[(K::V * K::T + K::V * K::D + 3) / 4 * 4];
    // two-pass affinity: items that survived the probe, and their valid-joint count so far
    unsigned short surv[K::V * K::T * K::D];
    unsigned char pcnt[K::V][K::T][K::D];
    signed char last[K::T];                  // ring index of the last pose
    signed char gv_n[K::T];
    signed char new_view[K::T];              // a matched camera is not in the track's view list yet
    signed char do_update[K::T];
    signed char out_row[K::T];               // output row of a reported track, -1 = not reported
    signed char life_flag[K::T];             // bit 0 = track kept, bit 1 = reported this frame
    signed char m[K::V];                     // detections per camera
    signed char conflict[K::V];              // camera needs the full assignment solver
    signed char um_n[K::V];
    unsigned char um_flag[K::V][K::D];
    signed char um[K::V][K::D];
    unsigned char nvj[K::T][K::J];
    unsigned char hyp_nviews[K::H];
    signed char hyp_cam[K::H][K::V];
    signed char hyp_det[K::H][K::V];
    unsigned char hyp_fail[K::H];
    signed char hyp_slot[K::H];
    PAM_HD signed char& t2d(int cam, int i) { return match[cam * K::T + i]; }
    PAM_HD signed char& d2t(int cam, int d) { return match[K::V * K::T + cam * K::D + d]; }
};

template <class K>
inline int arena_fixed_bytes() { return (int)((sizeof(SeqShared<K>) + 127) / 128 * 128); }
template <class K>
inline int cam_bytes() { return (int)((sizeof(CamShared<K>) + 127) / 128 * 128); }

// nbuf: detection buffers per sequence; raw_in_arena: keep the raw pose of the frame in the arena (latency
// oriented launches) instead of the sequence's HBM scratch (throughput oriented launches: 2-3 KB less)
inline void arena_layout(DevCfg& c, int nbuf, bool raw_in_arena) {
    c.caps = caps_class(c.V, c.D, c.J, c.max_trk);
    c.nbuf = nbuf;
    c.frame_floats = (c.V * c.D * c.J * 3 + 3) / 4 * 4;
    int o = c.caps == CAPS_SMALL ? arena_fixed_bytes<CapsSmall>() : (c.caps == CAPS_MID ? arena_fixed_bytes<CapsMid>() : arena_fixed_bytes<CapsMax>());
    c.a_dets = o;                                      // bulk-copy destination: 128-byte aligned
    o += nbuf * c.frame_floats * 4;
    o = (o + 15) / 16 * 16;
    c.a_raw = -1;
    if (raw_in_arena) { c.a_raw = o; o += 8 * c.max_trk * c.J * 3; }
    c.arena_bytes = (o + 127) / 128 * 128;
}
// re-lay the arena for a (larger) capacity class
inline void force_caps(DevCfg& c, int caps) {
    const int nbuf = c.nbuf;
    const bool raw = c.a_raw >= 0;
    arena_layout(c, nbuf, raw);
    if (caps > c.caps) {
        c.caps = caps;
        int o = c.caps == CAPS_MID ? arena_fixed_bytes<CapsMid>() : arena_fixed_bytes<CapsMax>();
        c.a_dets = o;
        o += nbuf * c.frame_floats * 4;
        o = (o + 15) / 16 * 16;
        c.a_raw = -1;
        if (raw) { c.a_raw = o; o += 8 * c.max_trk * c.J * 3; }
        c.arena_bytes = (o + 127) / 128 * 128;
    }
}
inline int cam_bytes_of(const DevCfg& c) {
    return c.caps == CAPS_SMALL ? cam_bytes<CapsSmall>() : (c.caps == CAPS_MID ? cam_bytes<CapsMid>() : cam_bytes<CapsMax>());
}

// Global-memory views of one sequence's state.
struct SeqGlobal {
    SeqHeader* hdr;
    TrkMeta* meta;
    double* hist;     // [max_trk][PAM_HIST][J][3]
    float* view;      // [max_trk][V][J][3]  (v, u, conf)
    float* vel;       // [max_trk][J][3]
    unsigned char* nv;  // [max_trk][J]
    double* init;       // new-track initialisation scratch
    double* raw;        // [max_trk][J][3]
    double* margin;     // [MG_COUNT]
    PAM_HD void bind(const DevCfg& c, char* base) {
        hdr = (SeqHeader*)(base + c.off_hdr);
        meta = (TrkMeta*)(base + c.off_meta);
        hist = (double*)(base + c.off_hist);
        view = (float*)(base + c.off_view);
        vel = (float*)(base + c.off_vel);
        nv = (unsigned char*)(base + c.off_nv);
        init = (double*)(base + c.off_init);
        raw = (double*)(base + c.off_raw);
        margin = (double*)(base + c.off_margin);
    }
};

// Per-frame outputs of one sequence (any pointer may be null).
struct FrameOut {
    int* count;            // [1]
    int* ids;              // [max_trk]
    float* joints;         // [max_trk][J][3]
    unsigned char* nviews; // [max_trk][J]
    int* assoc;            // [V][D]  track id matched to each detection, -1 = unmatched
    int* timing;           // [4]     SM cycles spent in association / update / initialisation / whole frame
    unsigned char* vlist;  // [max_trk][PAM_VLIST]  per reported track: usable views of this update, length of the track's
                           //         view list (= len(poses2d)), then the camera of every list entry in dict order,
                           //         bit 7 set when that view was matched this frame (ivclabpose.py:272-281)
};

// The working set of one sequence as the code sees it: fixed part, camera constants, run-time tail, global state.
// Lives in registers.
template <class K>
struct Seq {
    SeqShared<K>* sh;
    const CamShared<K>* cam;
    double* raw_;                            // raw pose of the frame in flight: arena tail or HBM scratch
    SeqGlobal g;
    PAM_HD const double* Pc(int k) const { return cam->P[k]; }
    PAM_HD const double* RKc(int k) const { return cam->RK[k]; }
    PAM_HD const double* posc(int k) const { return cam->pos[k]; }
    PAM_HD const double* Fc(int x, int y) const { return cam->F[x][y]; }
    PAM_HD void bind(const DevCfg& c, char* arena, const CamShared<K>* cams, char* state) {
        sh = (SeqShared<K>*)arena;
        cam = cams;
        g.bind(c, state);
        raw_ = c.a_raw >= 0 ? (double*)(arena + c.a_raw) : g.raw;
    }
    // new-track initialisation scratch (HBM: initialisation is rare)
    PAM_HD double* hyp_pose(const DevCfg&) const { return g.init; }
    PAM_HD double* hyp_cost(const DevCfg& c) const { return g.init + (int64_t)c.max_hyp * c.J * 3; }
    PAM_HD unsigned char* hyp_veto(const DevCfg& c) const { return (unsigned char*)(hyp_cost(c) + (int64_t)c.max_hyp * c.D); }
    PAM_HD unsigned char* hyp_nvj(const DevCfg& c) const { return hyp_veto(c) + (int64_t)c.max_hyp * c.D; }
};

struct HostCtx {
    static constexpr int kAffinityUnroll = 1;
    inline int tid() const { return 0; }
    inline int nthreads() const { return 1; }
    inline void sync() const {}
    inline void phase_sync() const {}
    inline void atomic_inc(int* p) const { *p += 1; }
    inline int atomic_inc_ret(int* p) const { return (*p)++; }
    inline void atomic_max_u64(unsigned long long* p, unsigned long long v) const { if (v > *p) *p = v; }
    inline void atomic_min_u64(unsigned long long* p, unsigned long long v) const { if (v < *p) *p = v; }
    inline void atomic_min(int* p, int v) const { if (v < *p) *p = v; }
    inline long long clock() const { return 0; }
    inline int enum_limit() const { return 1024; }      // as a one-warp group on the device
    static bool kTwoPassAffinity;            // set by the harness (PAM_HOSTEMU_TWOPASS=1)
};
struct NoHook {
    PAM_HD void dets_released() const {}
};

#if defined(PAM_MARGIN)
// smallest |x| seen (non-negative doubles order like their bit patterns)
template <class K>
PAM_HD void margin_note(const Seq<K>& sq, int k, double x) {
    x = fabs(x);
    if (!(x == x)) return;
#if defined(__CUDA_ARCH__)
    atomicMin((unsigned long long*)(sq.g.margin + k), (unsigned long long)__double_as_longlong(x));
#else
    if (x < sq.g.margin[k]) sq.g.margin[k] = x;
#endif
}
#define PAM_NOTE(k, x) margin_note(sq, k, x)
#else
#define PAM_NOTE(k, x) do { } while (0)
#endif

#define PAM_FOR(i, N) PAM_NOUNROLL for (int i = ctx.tid(), _n_##i = (N), _s_##i = ctx.nthreads(); i < _n_##i; i += _s_##i)
// Same, dealt from the highest thread downwards: small independent loops of a phase go to the threads
// (warps) that the phase's main loop leaves idle, so they run beside it instead of after it.
#define PAM_FOR_REV(i, N) \
    PAM_NOUNROLL for (int i = ctx.nthreads() - 1 - ctx.tid(), _n_##i = (N), _s_##i = ctx.nthreads(); i < _n_##i; i += _s_##i)

// camera constants: f32 in global memory (the reference's dtypes), widened once into shared memory by the
// whole CTA (t = thread index inside the CTA, nt = threads of the CTA).
struct CamConst {
    const float* P;      // [V][12]
    const float* RKinv;  // [V][9]
    const double* pos;   // [V][3]
    const float* F;      // [V][V][9]
};
template <class K>
PAM_HD void load_cameras(int t, int nt, const DevCfg& c, CamShared<K>* cam, const CamConst& cc) {
    const int V = c.V;
    PAM_NOUNROLL for (int i = t; i < V * 12; i += nt) cam->P[i / 12][i % 12] = (double)cc.P[i];
    PAM_NOUNROLL for (int i = t; i < V * 9; i += nt) cam->RK[i / 9][i % 9] = (double)cc.RKinv[i];
    PAM_NOUNROLL for (int i = t; i < V * 3; i += nt) cam->pos[i / 3][i % 3] = cc.pos[i];
    PAM_NOUNROLL for (int i = t; i < V * V * 9; i += nt) cam->F[i / (9 * V)][(i / 9) % V][i % 9] = (double)cc.F[i];
    PAM_NOUNROLL for (int i = t; i < V * V; i += nt) cam->F[i / V][i % V][9] = 0.0;      // padding, never read
}

template <class Ctx, class K>
PAM_HD void load_state(Ctx& ctx, const DevCfg& c, const Seq<K>& sq) {
    SeqShared<K>& sh = *sq.sh;
    const int* src = (const int*)sq.g.hdr;
    int* dst = (int*)&sh.hdr;
    PAM_FOR(i, (int)(sizeof(SeqHeader) / 4)) dst[i] = src[i];
    const int* srcm = (const int*)sq.g.meta;
    int* dstm = (int*)sh.trk;
    PAM_FOR(i, (int)(sizeof(TrkMeta) / 4) * c.max_trk) dstm[i] = srcm[i];
    PAM_FOR(i, (int)(sizeof(FrameScalars) / 4)) ((int*)&sh.fs)[i] = 0;
#if defined(PAM_MARGIN)
    // a fresh sequence (all-zero state) starts its margins at +inf
    PAM_FOR(i, MG_COUNT) if (src[3] == 0 && src[1] == 0) sq.g.margin[i] = HUGE_VAL;
#endif
}

template <class Ctx, class K>
PAM_HD void store_state(Ctx& ctx, const DevCfg& c, const Seq<K>& sq) {
    const SeqShared<K>& sh = *sq.sh;
    int* dst = (int*)sq.g.hdr;
    const int* src = (const int*)&sh.hdr;
    PAM_FOR(i, (int)(sizeof(SeqHeader) / 4)) dst[i] = src[i];
    int* dstm = (int*)sq.g.meta;
    const int* srcm = (const int*)sh.trk;
    PAM_FOR(i, (int)(sizeof(TrkMeta) / 4) * c.max_trk) dstm[i] = srcm[i];
}

// ------------------------------------------------------------------------------------------
// per-joint pieces of the update / init paths
// ------------------------------------------------------------------------------------------

// View accessors: the update path reads (u, v) of joint j straight from the ViewSrc records in the arena
// (no per-thread copies: with ~900 resident threads per SM thread-local arrays would spill to
// L2/HBM); the rare init path passes small local arrays.
struct SrcViews {
    const ViewSrc* vs;
    int j3;
    const double* w_age;
    PAM_HD int cid(int a) const { return vs[a].cid; }
    PAM_HD double u(int a) const { return (double)vs[a].p[j3 + 1]; }
    PAM_HD double v(int a) const { return (double)vs[a].p[j3]; }
    PAM_HD double w(int a) const { return w_age[vs[a].T]; }
    PAM_HD int age(int a) const { return vs[a].T; }
};
struct ArrayViews {
    const signed char* c;
    const double* uu;
    const double* vv;
    double w0;
    PAM_HD int cid(int a) const { return c[a]; }
    PAM_HD double u(int a) const { return uu[a]; }
    PAM_HD double v(int a) const { return vv[a]; }
    PAM_HD double w(int) const { return w0; }
    PAM_HD int age(int) const { return 0; }
};

// Weighted DLT over the surviving views (bit a of `alive`).
template <class K, class Views>
PAM_HD void dlt_from_views(const Seq<K>& sq, int Vt, const Views& vw, uint32_t alive, double* X) {
    // Gram / Cholesky fold first, stale views included with their weights e^{-lambda_t T}: as long as
    // the Cholesky pivots stay above 1e-6 of the diagonal (two fresh views, or one fresh view plus views
    // one frame old) the squared system resolves the solution to < 1e-10 relative.  Otherwise -- only
    // views two or three frames old next to at most one fresh view, or a noise-free system -- the rows
    // are folded again with Givens rotations, which keep the relative accuracy of the tiny rows.
    DltAccum acc;
    int path = -1;
    acc.reset(true);
    PAM_UNROLL_N(PAM_P5_UNROLL) for (int a = 0; a < Vt; ++a)
        if ((alive >> a) & 1u) acc.add_view(sq.Pc(vw.cid(a)), vw.u(a), vw.v(a), vw.w(a));
    acc.solve(X, &path);
    if (path < 0) {
        acc.reset(false);
        PAM_NOUNROLL for (int a = 0; a < Vt; ++a)
            if ((alive >> a) & 1u) acc.add_view(sq.Pc(vw.cid(a)), vw.u(a), vw.v(a), vw.w(a));
        acc.solve(X, &path);
    }
}

// Part-aware view filter + DLT for one joint of one track (update mode).
//   views 0..Vt-1 in the track's dict order; cid/T per view; (u, v) per view; next = predicted joint.
// Returns the number of surviving views; X = triangulated joint (or `next` when < 2 views).
template <class K>
PAM_HD int joint_update(const DevCfg& c, const Seq<K>& sq, int Vt, const SrcViews& vw, const double* next, double* X) {
    // conflict bit (a * 8 + b) for a < b: the pair's symmetric epipolar distance exceeds the threshold
    uint64_t conflict = 0ull;
    PAM_NOUNROLL for (int a = 0; a < Vt; ++a) {
        const double ua = vw.u(a), va = vw.v(a);
        const int ca = vw.cid(a);
        PAM_UNROLL_N(PAM_P5_UNROLL) for (int b = a + 1; b < Vt; ++b) {
            const double ub = vw.u(b), vb = vw.v(b);
            const int cb = vw.cid(b);
            double tab, nab, tba, nba;
            epi_raw_f64(sq.Fc(ca, cb), ua, va, ub, vb, tab, nab);
            epi_raw_f64(sq.Fc(cb, ca), ub, vb, ua, va, tba, nba);
#if !defined(PAM_MARGIN)
            // both one-way distances below the threshold => their mean is too => no conflict; decided
            // without a square root for the overwhelming majority of pairs
            if (tab * tab < c.joint_thr2 * nab && tba * tba < c.joint_thr2 * nba) continue;
#endif
            const double dab = fabs(tab) * ((nab == 0.0) ? 1.0 : rsqrt_f64(nab));
            const double dba = fabs(tba) * ((nba == 0.0) ? 1.0 : rsqrt_f64(nba));
            const double A = 1.0 - (dab + dba) / 2.0 * c.inv_joint_thr;
            PAM_NOTE(MG_JOINT_A, A);
            if (A < 0.0) conflict |= 1ull << (a * 8 + b);
        }
    }
    uint32_t alive = (1u << Vt) - 1u;
    if (conflict) {                // rare, so the ray distances are simply recomputed per conflict
        PAM_NOUNROLL for (int a = 0; a < Vt; ++a)
            PAM_NOUNROLL for (int b = a + 1; b < Vt; ++b) {
                if (!((conflict >> (a * 8 + b)) & 1ull)) continue;
                if (!((alive >> a) & 1u) || !((alive >> b) & 1u)) continue;
                const double ra = ray_point_distance(sq.RKc(vw.cid(a)), sq.posc(vw.cid(a)), vw.u(a), vw.v(a), next);
                const double rb = ray_point_distance(sq.RKc(vw.cid(b)), sq.posc(vw.cid(b)), vw.u(b), vw.v(b), next);
                PAM_NOTE(MG_RAY, (ra - rb) / (ra > rb ? ra : rb));
                if (ra > rb) alive &= ~(1u << a); else alive &= ~(1u << b);
            }
    }
    const int nv = popcount32(alive);
    if (nv < 2) {
        X[0] = next[0]; X[1] = next[1]; X[2] = next[2];
        return nv;
    }
    dlt_from_views(sq, Vt, vw, alive, X);
    return nv;
}

// Same for a hypothesis (init mode, float32 affinities, row-sum rule).  Returns surviving views.
template <class K>
PAM_HD int joint_init(const DevCfg& c, const Seq<K>& sq, int Vt, const signed char* cid, const double* u, const double* v,
                      double* X) {
    float A[PAM_MAX_V][PAM_MAX_V];
    bool any = false;
    PAM_NOUNROLL for (int a = 0; a < Vt; ++a) {
        A[a][a] = 1.0f - 0.0f / c.init_thr_f32;
        PAM_NOUNROLL for (int b = a + 1; b < Vt; ++b) {
            double d1, d2;
            epi_pair_cv(sq.Fc(cid[a], cid[b]), u[a], v[a], u[b], v[b], d1, d2);
            float Df = (float)((d1 + d2) / 2.0);
            float Af = 1.0f - Df / c.init_thr_f32;
            A[a][b] = Af; A[b][a] = Af;
            PAM_NOTE(MG_INIT_A, (double)Af);
            if (Af < 0.0f) any = true;
        }
    }
    uint32_t alive = (1u << Vt) - 1u;
    if (any) {
        PAM_NOUNROLL for (int a = 0; a < Vt; ++a)
            PAM_NOUNROLL for (int b = a + 1; b < Vt; ++b) {
                if (!(A[a][b] < 0.0f)) continue;
                if (!((alive >> a) & 1u) || !((alive >> b) & 1u)) continue;
                float s1 = np_sum(A[a], Vt), s2 = np_sum(A[b], Vt);
                PAM_NOTE(MG_INIT_ROWSUM, (double)(s1 - s2));
                if (s1 > s2) alive &= ~(1u << b); else alive &= ~(1u << a);
            }
    }
    const int nv = popcount32(alive);
    if (nv < 2) return nv;
    dlt_from_views(sq, Vt, ArrayViews{cid, u, v, c.w_age[0]}, alive, X);
    return nv;
}

// Hypothesis.calculate_cost for hypothesis h against detection (cam c2, pose o) of this frame.
template <class K>
PAM_HD double hyp_cost(const DevCfg& c, const Seq<K>& sq, const float* dets, int h, int c2, int d2, bool& veto) {
    const SeqShared<K>& sh = *sq.sh;
    const int J = c.J;
    const float* o = dets + ((int64_t)(c2 * c.D + d2) * J) * 3;
    double total = 0.0;
    veto = false;
    const int nvw = sh.hyp_nviews[h];
    const signed char* hcam = sh.hyp_cam[h];
    const signed char* hdet = sh.hyp_det[h];
    PAM_NOUNROLL for (int k = 0; k < nvw; ++k) {
        const int c1 = hcam[k];
        const float* p = dets + ((int64_t)(c1 * c.D + hdet[k]) * J) * 3;
        NpSumStream<double> acc;
        acc.begin(J);
        PAM_NOUNROLL for (int j = 0; j < J; ++j) {
            double d1, d2v;
            epi_pair_cv(sq.Fc(c1, c2), (double)p[j * 3 + 1], (double)p[j * 3 + 0], (double)o[j * 3 + 1],
                        (double)o[j * 3 + 0], d1, d2v);
            acc.push((d1 * (double)p[j * 3 + 2] + d2v * (double)o[j * 3 + 2]) / 2.0);
        }
        double pc = acc.total() / (double)J / c.epi_thr;
        total += pc;
        if (sh.believe[c2][d2] > c.veto_believe) PAM_NOTE(MG_VETO, pc - 1.0);
        if (pc > 1.0 && sh.believe[c2][d2] > c.veto_believe) veto = true;
    }
    return total / (double)nvw;
}

// ------------------------------------------------------------------------------------------
// the frame
// ------------------------------------------------------------------------------------------
// per-track success test of update_3dpose (IterativeTracker.py:324-325, 369), valid after phase 5
template <class K>
PAM_HD bool track_ok(const DevCfg& c, const Seq<K>& sq, int i) {
    const SeqShared<K>& sh = *sq.sh;
    return sh.do_update[i] && !((double)sh.fail[i] > c.fail_limit) &&
           sh.trk[sh.hdr.order[i]].hist_len < PAM_HIST;
}
// will track i be reported this frame (Confirmed after this update, ivclabpose.py:266)?
template <class K>
PAM_HD bool track_reported(const DevCfg& c, const Seq<K>& sq, int i) {
    const SeqShared<K>& sh = *sq.sh;
    if (!track_ok(c, sq, i)) return false;
    const TrkMeta& t = sh.trk[sh.hdr.order[i]];
    return t.state == ST_CONFIRMED || (t.state == ST_TENTATIVE && t.hits + 1 >= c.n_init);
}

// `dets`/`counts`: this frame's detections (staged in the arena by the kernel).  `gin`: the launch's
// detection tensor of this sequence in global memory, frame id `gin_frame0` at offset 0: a view that
// was matched 1-3 frames ago is read from there instead of being copied into the track state every
// frame; persist_views() writes the views of the launch into the state once, at the end.
// `hook.dets_released()` is called by every thread once the staged detections are no longer needed
// (single-buffered launches start the copy of the next frame there).
// The caller must synchronise the group after frame_step returns.
// two assignments whose totals agree to this relative width are left to the general solver (2^-40)
#define PAM_ASSIGN_TIE 9.094947017729282e-13

// index of the k-th set bit of m (k < popcount(m))
PAM_HD int nth_set_bit(uint32_t m, int k) {
    PAM_NOUNROLL for (; k > 0; --k) m &= m - 1u;
    return ctz32(m);
}

// One camera's assignment problem (IterativeTracker.py:150-160: scipy's linear_sum_assignment on -affinity, pairs with
// affinity <= 0 dropped afterwards), solved by the whole thread group when it is small.  Entries that are not
// positive contribute nothing and any matching of positive entries extends to a complete assignment, so the optimum
// is the maximum-weight matching of the positive entries.  Pairs already fixed by the quick test are isolated; the
// rest -- r rows and c columns that hold a positive entry -- is enumerated: every injective map of the shorter side
// into the longer one is a candidate (p! / (p - o)! of them), its total is summed in index order of the shorter side
// (candidates with the same positive pairs have bit-identical totals), the threads keep their best candidate and two
// shared-memory atomics pick the largest total.  The solver of the reference reaches the same optimum unless two
// different matchings tie to within rounding.  Returns false (nothing written) when the problem has more than `limit`
// candidates or when two different matchings tie to within 2^-40: the caller then runs the general solver.
template <class Ctx, class K>
PAM_HD bool assign_enumerated(Ctx& ctx, SeqShared<K>& sh, int cam, int n, int mm, int limit) {
    const double (*A)[K::D] = sh.aff[cam];
    FrameScalars& fs = sh.fs;
    uint32_t rows = 0u, cols = 0u;
    PAM_NOUNROLL for (int i = 0; i < n; ++i) {
        if (sh.t2d(cam, i) >= 0) continue;                      // fixed by the quick test
        PAM_NOUNROLL for (int d = 0; d < mm; ++d)
            if (A[i][d] > 0.0) { rows |= 1u << i; cols |= 1u << d; }
    }
    const int r = popcount32(rows), cc = popcount32(cols);
    const bool by_rows = r <= cc;                               // the shorter side is mapped into the longer one
    const int o = by_rows ? r : cc, p = by_rows ? cc : r;
    const uint32_t outer = by_rows ? rows : cols, pool = by_rows ? cols : rows;
    int P = 1;
    PAM_NOUNROLL for (int q = 0; q < o; ++q) {
        P *= p - q;
        if (P > limit) return false;                            // uniform: every thread sees the same matrix
    }
    ctx.sync();                                                 // the previous camera's verdict has been read by everyone
    if (ctx.tid() == 0) { fs.bf_best = 0ull; fs.bf_k = P; fs.bf_sig_lo = ~0ull; fs.bf_sig_hi = 0ull; }
    ctx.sync();
    // total of candidate k and a signature of its positive pairs (order-independent 64-bit mix)
    auto total = [&](int k, unsigned long long& sig) {
        uint32_t avail = pool, left = outer;
        double sum = 0.0;
        sig = 0ull;
        PAM_NOUNROLL for (int q = 0; q < o; ++q) {
            const int base = p - q, digit = k % base;
            k /= base;
            const int a = ctz32(left), b = nth_set_bit(avail, digit);
            left &= left - 1u;
            avail &= ~(1u << b);
            const double x = by_rows ? A[a][b] : A[b][a];
            sum += x;
            if (x > 0.0) {
                unsigned long long z = (unsigned long long)((by_rows ? a : b) * 64 + (by_rows ? b : a) + 1) * 0x9E3779B97F4A7C15ull;
                z ^= z >> 31; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 29;
                sig ^= z;
            }
        }
        return sum;
    };
    auto bits_of = [](double x) { unsigned long long b; memcpy(&b, &x, 8); return b; };   // non-negative: ordered like x
    double best = -1.0;
    PAM_FOR(k, P) {
        unsigned long long sig;
        const double t = total(k, sig);
        if (t > best) best = t;
    }
    if (best >= 0.0) ctx.atomic_max_u64(&fs.bf_best, bits_of(best));
    ctx.sync();
    // the candidates within 2^-40 of the maximum (the totals of two matchings that tie in exact arithmetic can differ in
    // the last bits, their terms being added in different orders): lowest index among those that reach it, and whether
    // they all hold the same positive pairs
    {
        double top;
        const unsigned long long tb = fs.bf_best;
        memcpy(&top, &tb, 8);
        const double near_top = top * (1.0 - PAM_ASSIGN_TIE);
        if (best >= near_top) {
            PAM_FOR(k, P) {
                unsigned long long sig;
                const double t = total(k, sig);
                if (!(t >= near_top)) continue;
                if (bits_of(t) == tb) ctx.atomic_min(&fs.bf_k, k);
                ctx.atomic_min_u64(&fs.bf_sig_lo, sig);
                ctx.atomic_max_u64(&fs.bf_sig_hi, sig);
            }
        }
    }
    ctx.sync();
    // two different matchings with (nearly) the same total -- duplicated detections or tracks: which one the
    // reference's solver returns depends on its pivoting order and its own rounding, so that solver decides
    if (fs.bf_sig_lo != fs.bf_sig_hi) return false;
    if (ctx.tid() == 0) {
        int k = fs.bf_k;
        uint32_t avail = pool, left = outer;
        PAM_NOUNROLL for (int q = 0; q < o; ++q) {
            const int base = p - q, digit = k % base;
            k /= base;
            const int a = ctz32(left), b = nth_set_bit(avail, digit);
            left &= left - 1u;
            avail &= ~(1u << b);
            const int i = by_rows ? a : b, d = by_rows ? b : a;
            if (A[i][d] > 0.0) { sh.t2d(cam, i) = (signed char)d; sh.d2t(cam, d) = (signed char)i; }
        }
    }
    return true;
}

// The same problem when, besides the pairs fixed by the quick test, at most two tracks and two detections -- or one
// track / one detection against several -- hold positive entries: the rule when two people cross.  One thread
// decides it in closed form (every camera in parallel).  Returns false when the camera needs more than that, or when
// the alternatives tie to within 2^-40.
template <class K>
PAM_HD bool assign_closed_form(SeqShared<K>& sh, int cam, int n, int mm) {
    const double (*A)[K::D] = sh.aff[cam];
    uint32_t rows = 0u, cols = 0u;
    PAM_NOUNROLL for (int i = 0; i < n; ++i) {
        if (sh.t2d(cam, i) >= 0) continue;
        PAM_NOUNROLL for (int d = 0; d < mm; ++d)
            if (A[i][d] > 0.0) { rows |= 1u << i; cols |= 1u << d; }
    }
    const int r = popcount32(rows), cc = popcount32(cols);
    auto take = [&](int i, int d) {
        if (A[i][d] > 0.0) { sh.t2d(cam, i) = (signed char)d; sh.d2t(cam, d) = (signed char)i; }
    };
    if (r == 1 || cc == 1) {
        // one track with several candidate detections, or several tracks after one detection: the largest entry
        const bool one_row = r == 1;
        const int fixed = ctz32(one_row ? rows : cols);
        uint32_t rest = one_row ? cols : rows;
        double best = 0.0, second = 0.0;
        int arg = -1;
        PAM_NOUNROLL for (; rest; rest &= rest - 1u) {
            const int x = ctz32(rest);
            const double a = one_row ? A[fixed][x] : A[x][fixed];
            if (a > best) { second = best; best = a; arg = x; }
            else if (a > second) second = a;
        }
        if (arg < 0 || second >= best * (1.0 - PAM_ASSIGN_TIE)) return false;     // (near) tie: the general solver decides
        if (one_row) take(fixed, arg); else take(arg, fixed);
        return true;
    }
    if (r == 2 && cc == 2) {
        const int i0 = ctz32(rows), i1 = ctz32(rows & (rows - 1u)), d0 = ctz32(cols), d1 = ctz32(cols & (cols - 1u));
        const double straight = A[i0][d0] + A[i1][d1], crossed = A[i0][d1] + A[i1][d0];
        const double hi = straight > crossed ? straight : crossed, lo = straight > crossed ? crossed : straight;
        if (lo >= hi * (1.0 - PAM_ASSIGN_TIE)) return false;                      // (near) tie: the general solver decides
        if (straight > crossed) { take(i0, d0); take(i1, d1); }
        else { take(i0, d1); take(i1, d0); }
        return true;
    }
    return false;
}

// relative half-width of the guard band of the sign-only affinity test (phase 2); tests widen it to drive
// joints through the exact fallback
#if !defined(PAM_SIGN_BAND)
#define PAM_SIGN_BAND 5.9604644775390625e-08     /* 2^-24 */
#endif

template <class Ctx, class K, class Hook>
PAM_HD void frame_step(Ctx& ctx, const DevCfg& c, const Seq<K>& sq, int frame, const float* dets /* [V][D][J][3] */,
                       const int* counts /* [V] */, const FrameOut& out, const float* gin, int gin_frame0, Hook& hook) {
    const int V = c.V, J = c.J, D = c.D, MT = c.max_trk;
    const int J3 = J * 3;
    SeqShared<K>& sh = *sq.sh;
    SeqHeader& hdr = sh.hdr;
    FrameScalars& fs = sh.fs;
    TrkMeta* const trk = sh.trk;
    const SeqGlobal& g = sq.g;
    if (hdr.status != SEQ_OK) {   // uniform: status only changes between syncs
        if (ctx.tid() == 0 && out.count) *out.count = 0;
        hook.dets_released();
        for (int k = 0; k < 7; ++k) ctx.phase_sync();     // keep in step with the other sequences of a convoy
        return;
    }
    const int n = hdr.ntracks;
    long long tk0 = 0, tk1 = 0, tk2 = 0;
    if (out.timing) tk0 = ctx.clock();

    // ---- phase 1: ageing + snapshot (IterativeTracker.py:126-129) ---------------------------------
    PAM_FOR(i, n) {
        TrkMeta& t = trk[hdr.order[i]];
        t.already = 0; t.age += 1; t.tsu += 1;
        const int last = (t.hist_start + t.hist_len - 1) % PAM_HIST;
        const int dt = frame - t.hist_time[last];
        sh.last[i] = (signed char)last;
        sh.dt[i] = dt;
        if (dt >= 0 && dt < 16) {
            sh.inv_denom[i] = c.inv_denom_tab[dt];                 // IterativeTracker.py:143
            sh.inv_decay[i] = c.inv_decay_tab[dt];                 // IterativeTracker.py:148
        } else {
            sh.inv_denom[i] = 1.0 / (c.alpha2d * (double)dt);
            sh.inv_decay[i] = 1.0 / exp(c.lambda_a * (double)dt);
        }
        {
            // sign-only affinity (phase 2): valid when dt > 0 and the decay factor is an ordinary positive number, so
            // that "more than min_valid joints with c > 0" is the same as "affinity > 0"; otherwise lo > hi sends every
            // joint of this track through the exact evaluation
            const double th = c.alpha2d * (double)dt, th2 = th * th;
            const bool plain = dt > 0 && th > 0.0 && th2 < 1e300 && sh.inv_decay[i] > 1e-150 && sh.inv_decay[i] < 1e150;
            sh.thr2_lo[i] = plain ? th2 * (1.0 - PAM_SIGN_BAND) : -1.0;
            sh.thr2_hi[i] = plain ? th2 * (1.0 + PAM_SIGN_BAND) : HUGE_VAL;
        }
        sh.fail[i] = 0;
        sh.new_view[i] = 0;
    }
    if (ctx.tid() == ctx.nthreads() - 1) { fs.any_conflict = 0; fs.any_deleted = 0; fs.n_surv = 0; fs.unsure = 0; }
    PAM_FOR_REV(cc, V) {
        int mm = counts[cc];
        if (mm > D || mm < 0) { hdr.warn |= WARN_DET_OVERFLOW; fs.warned = 1; mm = (mm < 0) ? 0 : D; }
        sh.m[cc] = (signed char)mm;
        sh.conflict[cc] = 0;
    }
    {
        int* mw = (int*)sh.match;                     // t2d and d2t, contiguous
        PAM_FOR_REV(i, (int)(sizeof(sh.match) / 4)) mw[i] = -1;
    }
    ctx.phase_sync();

    // ---- phase 2: track x detection affinity (IterativeTracker.py:139-149) ----------------------
    // joints j0..j1-1 of one (track, camera, detection): reprojection of the track's last pose into the camera
    // (ivclabpose.py:91-98; recomputed per detection: cheaper than staging V x n x J pixel pairs in shared memory),
    // distance to the detection, c = 1 - d / (alpha2d dt); running sum and count of the joints with c > 0
    auto affinity_joints = [&](int i, int cam, int d, int j0, int j1, double& sum, int& cnt) {
        const double* X = g.hist + (int64_t)(hdr.order[i] * PAM_HIST + sh.last[i]) * J3;
        const double* P = sq.Pc(cam);
        const double p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
        const double p8 = P[8], p9 = P[9], p10 = P[10], p11 = P[11];
        const float* q = dets + (int64_t)(cam * D + d) * J3;
        const double inv_denom = sh.inv_denom[i];
        // joints in flight per thread (the chain per joint is ~20 deep): 2 in the register-lean variants,
        // 4 where the launch shape leaves the full register budget
        PAM_UNROLL_N(Ctx::kAffinityUnroll) for (int j = j0; j < j1; ++j) {
            const double x = X[j * 3], y = X[j * 3 + 1], z = X[j * 3 + 2];
            const double iw = rcp_f64(p8 * x + p9 * y + p10 * z + p11);
            const double dv = (p4 * x + p5 * y + p6 * z + p7) * iw - (double)q[j * 3 + 0];
            const double du = (p0 * x + p1 * y + p2 * z + p3) * iw - (double)q[j * 3 + 1];
            double cj = 1.0 - sqrt_f64(dv * dv + du * du) * inv_denom;
            PAM_NOTE(MG_ASSOC_C, cj);
            if (cj > 0.0) { sum += cj; ++cnt; }
        }
    };
    auto affinity_value = [&](int i, double sum, int cnt) {
        double a = (cnt > c.min_valid) ? sum * rcp_f64((double)cnt) : 0.0;
        a = a * sh.inv_decay[i];
        if (a != a) a = 0.0;
        if (a > 0.0) PAM_NOTE(MG_ASSIGN, a);
        return a;
    };
#if !defined(PAM_MARGIN)
    // Sign-only evaluation.  Phase 3 needs the VALUE of an affinity only in a camera whose positive entries do not
    // already form a matching; everywhere else it needs the sign, and affinity > 0 <=> more than min_valid joints
    // have c > 0 <=> |x_proj - x_det| < alpha2d dt.  That comparison is made without the division and the square
    // root, on |n - q w|^2 against (alpha2d dt)^2 w^2 with a relative guard band of 2^-24 on either side (the two
    // forms differ by ~1e-13 relative at most).  A joint inside the band -- or any joint of a track whose dt or
    // decay is not an ordinary number -- makes the whole frame fall back to the exact expression.  Otherwise the entry
    // holds 1 (positive) or 0; the cameras that need values get them in phase 3.
    auto count_joints = [&](int i, int cam, int d, int j0, int j1, int& cnt) {
        const double* X = g.hist + (int64_t)(hdr.order[i] * PAM_HIST + sh.last[i]) * J3;
        const double* P = sq.Pc(cam);
        const double p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
        const double p8 = P[8], p9 = P[9], p10 = P[10], p11 = P[11];
        const float* q = dets + (int64_t)(cam * D + d) * J3;
        const double lo = sh.thr2_lo[i], hi = sh.thr2_hi[i];
        uint32_t unsure = 0u;                        // joints inside the guard band (J <= 32)
        PAM_UNROLL_N(Ctx::kAffinityUnroll) for (int j = j0; j < j1; ++j) {      // straight-line body: the joints interleave
            const double x = X[j * 3], y = X[j * 3 + 1], z = X[j * 3 + 2];
            const double w = p8 * x + p9 * y + p10 * z + p11;
            const double ev = (p4 * x + p5 * y + p6 * z + p7) - (double)q[j * 3 + 0] * w;
            const double eu = (p0 * x + p1 * y + p2 * z + p3) - (double)q[j * 3 + 1] * w;
            const double e2 = ev * ev + eu * eu, w2 = w * w;
            const bool inside = e2 < lo * w2, outside = e2 > hi * w2;
            cnt += inside ? 1 : 0;
            unsure |= (inside || outside) ? 0u : (1u << j);
        }
        if (unsure) fs.unsure = 1;                   // practically never
    };
    constexpr bool kSignOnly = true;
#else
    auto count_joints = [&](int, int, int, int, int, int&) {};
    constexpr bool kSignOnly = false;     // the margin build looks at every c
#endif
    if (Ctx::kTwoPassAffinity && c.aff_probe > 0 && V * n * D > ctx.nthreads()) {      // pays from two sweeps on
        // Throughput launches.  Most (track, detection) pairs belong to different people: after the first
        // `aff_probe` joints such a pair can no longer collect more than min_valid joints with c > 0, so its
        // affinity is exactly 0 (IterativeTracker.py:145-147) and it is dropped.  The survivors -- about one
        // per (track, camera) -- are compacted and finished in a second, densely packed pass; their partial
        // sums continue in the same order, so every affinity is bit-identical to the single-pass value.
        const int JA = c.aff_probe;
        PAM_FOR(it, V * n * D) {
            const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
            if (d >= sh.m[cam]) continue;
            double sum = 0.0;
            int cnt = 0;
            if (kSignOnly) count_joints(i, cam, d, 0, JA, cnt);
            else affinity_joints(i, cam, d, 0, JA, sum, cnt);
            if (cnt + (J - JA) <= c.min_valid) { sh.aff[cam][i][d] = 0.0; continue; }
            sh.aff[cam][i][d] = sum;
            sh.pcnt[cam][i][d] = (unsigned char)cnt;
            sh.surv[ctx.atomic_inc_ret(&fs.n_surv)] = (unsigned short)it;
        }
        ctx.sync();
        PAM_FOR(k, fs.n_surv) {
            const int it = sh.surv[k];
            const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
            double sum = sh.aff[cam][i][d];
            int cnt = sh.pcnt[cam][i][d];
            if (kSignOnly) {
                count_joints(i, cam, d, JA, J, cnt);
                sh.aff[cam][i][d] = (cnt > c.min_valid) ? 1.0 : 0.0;
            } else {
                affinity_joints(i, cam, d, JA, J, sum, cnt);
                sh.aff[cam][i][d] = affinity_value(i, sum, cnt);
            }
        }
    } else {
        PAM_FOR(it, V * n * D) {
            const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
            if (d >= sh.m[cam]) continue;
            double sum = 0.0;
            int cnt = 0;
            if (kSignOnly) {
                count_joints(i, cam, d, 0, J, cnt);
                sh.aff[cam][i][d] = (cnt > c.min_valid) ? 1.0 : 0.0;
            } else {
                affinity_joints(i, cam, d, 0, J, sum, cnt);
                sh.aff[cam][i][d] = affinity_value(i, sum, cnt);
            }
        }
    }
    ctx.phase_sync();
    // every affinity of the frame (all = true), or the positive entries of the cameras that need the solver, by the
    // exact expression
    auto exact_affinities = [&](bool all) {
        if (!all) {
            // few entries, on the critical path of the CTA's convoy: one joint per thread, the terms parked in the
            // (idle) raw-pose buffer, then summed in joint order by one thread per entry -- the same additions in the
            // same order as the loop below
            if (ctx.tid() == 0) fs.n_surv = 0;
            ctx.sync();
            PAM_FOR(it, V * n * D) {
                const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
                if (d < sh.m[cam] && sh.conflict[cam] && sh.aff[cam][i][d] > 0.0)
                    sh.surv[ctx.atomic_inc_ret(&fs.n_surv)] = (unsigned short)it;
            }
            ctx.sync();
            const int E = fs.n_surv;
            if (E <= c.max_trk * 3) {                          // the buffer holds max_trk * J * 3 doubles
                double* const term = sq.raw_;
                PAM_FOR(x, E * J) {
                    const int e = fast_div(x, c.inv_J), j = x - e * J, it = sh.surv[e];
                    const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
                    double cj = 0.0;
                    int one = 0;
                    affinity_joints(i, cam, d, j, j + 1, cj, one);
                    term[x] = cj;                              // c_j where it is positive, else 0
                }
                ctx.sync();
                PAM_FOR(e, E) {
                    const int it = sh.surv[e];
                    const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
                    double sum = 0.0;
                    int cnt = 0;
                    PAM_NOUNROLL for (int j = 0; j < J; ++j) {
                        const double cj = term[e * J + j];
                        if (cj > 0.0) { sum += cj; ++cnt; }
                    }
                    sh.aff[cam][i][d] = affinity_value(i, sum, cnt);
                }
                ctx.sync();
                return;
            }
        }
        PAM_FOR(it, V * n * D) {
            const int i = fast_div(it, c.inv_VD), rem = it - i * (V * D), cam = fast_div(rem, c.inv_D), d = rem - cam * D;
            if (d >= sh.m[cam]) continue;
            if (!all && (!sh.conflict[cam] || !(sh.aff[cam][i][d] > 0.0))) continue;
            double sum = 0.0;
            int cnt = 0;
            affinity_joints(i, cam, d, 0, J, sum, cnt);
            sh.aff[cam][i][d] = affinity_value(i, sum, cnt);
        }
        ctx.sync();
    };
    if (kSignOnly && fs.unsure) exact_affinities(true);      // uniform

    // ---- phase 3: one assignment problem per camera (IterativeTracker.py:150-160) ---------------
    // Only pairs with affinity > 0 are ever accepted.  When the positive entries of a camera's
    // matrix already form a matching (at most one per row and per column) every optimal assignment
    // contains exactly those pairs, so they are taken directly; otherwise the camera is flagged
    // and solved with the full shortest-augmenting-path algorithm below.
    // The thread that finds the match of (track, camera) applies it to the track's view list at once
    // (add_pose, IterativeTracker.py:289-298): a pair that is the only positive entry of its row and of
    // its column belongs to every optimal assignment, so this stays valid if the camera is re-solved below.
    // A camera that is not in the list yet is appended by the per-track pass (insertion order = camera order).
    auto apply_match = [&](int i, int cam, int d) {
        TrkMeta& t = trk[hdr.order[i]];
        const int k = t.view_slot[cam];               // view slot of this camera, -1 = not in the dict yet
        if (k >= 0) { t.view_time[k] = frame; t.view_det[k] = (signed char)d; }
        else sh.new_view[i] = 1;
        t.already = 1;
    };
    PAM_FOR(it, V * n) {
        const int i = fast_div(it, c.inv_V), cam = it - i * V;
        const int mm = sh.m[cam];
        const double (*A)[K::D] = sh.aff[cam];
        int cnt = 0, arg = -1;
        PAM_UNROLL4 for (int d = 0; d < mm; ++d)           // independent loads: four in flight
            if (A[i][d] > 0.0) { ++cnt; arg = d; }
        if (cnt == 1) {
            int col = 0;
            PAM_UNROLL4 for (int k = 0; k < n; ++k) col += (A[k][arg] > 0.0) ? 1 : 0;
            if (col == 1) {
                sh.t2d(cam, i) = (signed char)arg; sh.d2t(cam, arg) = (signed char)i;
                apply_match(i, cam, arg);
            } else { sh.conflict[cam] = 1; fs.any_conflict = 1; }
        } else if (cnt > 1) {
            sh.conflict[cam] = 1; fs.any_conflict = 1;
        }
    }
    ctx.phase_sync();
    if (fs.any_conflict) {   // uniform
        // the solver compares affinities: the positive entries of the flagged cameras get their values now
        if (kSignOnly) exact_affinities(false);
        // the quick test marked these cameras 1.  Two tracks and two detections that cross, or one against several, are
        // decided in closed form by one thread per camera (-> 3); other small problems are enumerated by the whole
        // group; the ones left for the general solver become 2
        PAM_FOR(cam, V)
            if (sh.conflict[cam] && assign_closed_form(sh, cam, n, sh.m[cam])) sh.conflict[cam] = 3;
        ctx.sync();
        PAM_NOUNROLL for (int cam = 0; cam < V; ++cam) {         // uniform
            if (sh.conflict[cam] != 1) continue;
            const bool done = assign_enumerated(ctx, sh, cam, n, sh.m[cam], ctx.enum_limit());
            if (!done && ctx.tid() == 0) sh.conflict[cam] = 2;
        }
        ctx.sync();
        PAM_FOR(cam, V) {
            if (sh.conflict[cam] != 2) continue;
            const int mm = sh.m[cam];
            const double (*A)[K::D] = sh.aff[cam];
            signed char* t2d = &sh.t2d(cam, 0);
            signed char* d2t = &sh.d2t(cam, 0);
            PAM_NOUNROLL for (int i = 0; i < n; ++i) t2d[i] = -1;
            PAM_NOUNROLL for (int d = 0; d < D; ++d) d2t[d] = -1;
            int col4row[PAM_LSAP_N];
            lsap_solve<PAM_LSAP_N>(n, mm, [&](int i, int d) { return -A[i][d]; }, col4row);
            PAM_NOUNROLL for (int i = 0; i < n; ++i) {
                int d = col4row[i];
                if (d >= 0 && A[i][d] > 0.0) { t2d[i] = (signed char)d; d2t[d] = (signed char)i; }
            }
        }
        ctx.sync();
        PAM_FOR(it, V * n) {
            const int i = fast_div(it, c.inv_V), cam = it - i * V;
            if (sh.conflict[cam] && sh.t2d(cam, i) >= 0) apply_match(i, cam, sh.t2d(cam, i));
        }
        ctx.sync();
    }
    if (out.timing) tk1 = ctx.clock();

    // ---- phase 4: gather the usable views of every track in dict-insertion order
    //      (IterativeTracker.py:310-325); mean confidence of every unmatched detection (calculate.py:8-14)
    PAM_FOR(it, n * V) {
        const int i = fast_div(it, c.inv_V), k = it - i * V;
        const int s = hdr.order[i];
        TrkMeta& t = trk[s];
        const int nw = sh.new_view[i];
        if (nw) {
            // rare (a camera sees this track for the first time): one thread appends the new cameras in
            // camera order and gathers the whole list
            if (k != 0) continue;
            PAM_NOUNROLL for (int cam = 0; cam < V; ++cam) {
                if (sh.t2d(cam, i) < 0 || t.view_slot[cam] >= 0) continue;
                const int kk = t.nviews++;
                t.view_cid[kk] = cam; t.view_slot[cam] = (signed char)kk;
                t.view_time[kk] = frame; t.view_det[kk] = sh.t2d(cam, i);
            }
        } else if (k > 0 && k >= t.nviews) {
            continue;
        }
        // view k (all views when this thread gathers alone): its place = number of usable views before it
        const int k0 = nw ? 0 : k, k1 = nw ? t.nviews : k + 1;
        int cnt = 0;
        if (t.already) {
            PAM_NOUNROLL for (int kk = 0; kk < k0; ++kk) cnt += (frame - t.view_time[kk] <= c.stale_window) ? 1 : 0;
            PAM_NOUNROLL for (int kk = k0; kk < k1; ++kk) {
                const int age = frame - t.view_time[kk];
                if (age > c.stale_window) continue;
                const int cam = t.view_cid[kk];
                ViewSrc& vs = sh.vsrc[i][cnt++];
                vs.cid = cam;
                vs.T = age;
                // a view matched this frame is read straight from the staged detections
                const int tl = t.view_time[kk] - gin_frame0;   // frame index inside this launch (< 0: earlier launch)
                vs.p = (age == 0) ? dets + (int64_t)(cam * D + sh.t2d(cam, i)) * J3
                     : (tl >= 0) ? gin + ((int64_t)tl * V * D + cam * D + t.view_det[kk]) * J3
                                 : g.view + (int64_t)(s * V + kk) * J3;
            }
        }
        if (k1 >= t.nviews) {                        // the thread of the last view closes the list
            sh.gv_n[i] = (signed char)cnt;
            sh.do_update[i] = (t.already && cnt >= 2) ? 1 : 0;
        }
    }
    PAM_FOR_REV(it, V * D) {
        const int cam = fast_div(it, c.inv_D), d = it - cam * D;
        sh.um_flag[cam][d] = 0;
        const int i = (d < sh.m[cam]) ? sh.d2t(cam, d) : -1;
        if (out.assoc) out.assoc[it] = (i >= 0) ? trk[hdr.order[i]].track_id : -1;
        if (d >= sh.m[cam] || i >= 0) continue;   // the mean confidence only matters for unmatched detections
        const float* q = dets + (int64_t)(cam * D + d) * J3;
        const double b = mean_confidence(q, J);
        PAM_NOTE(MG_BELIEVE, b - c.conf_thr);
        sh.believe[cam][d] = b;
        sh.um_flag[cam][d] = (b > c.conf_thr) ? 1 : 0;
    }
    ctx.phase_sync();

    // ---- phase 5: per (track, joint): part-aware view filter + DLT (IterativeTracker.py:337-369);
    //      unmatched lists ---------------------------------------------------------------------
    {
        double* const raw = sq.raw_;
        PAM_FOR(it, n * J) {
            const int i = fast_div(it, c.inv_J), j = it - i * J;
            if (!sh.do_update[i]) continue;
            const int s = hdr.order[i];
            const double* Xl = g.hist + ((int64_t)(s * PAM_HIST + sh.last[i]) * J + j) * 3;
            const float* vel = g.vel + (int64_t)(s * J + j) * 3;
            const float fdt = (float)sh.dt[i];
            double next[3];
            for (int k = 0; k < 3; ++k) next[k] = Xl[k] + (double)(vel[k] * fdt);
            const int Vt = sh.gv_n[i];
            double X[3];
            const int nv = joint_update(c, sq, Vt, SrcViews{sh.vsrc[i], j * 3, c.w_age}, next, X);
            sh.nvj[i][j] = (unsigned char)nv;
            if (nv < 2) ctx.atomic_inc(&sh.fail[i]);
            double* r = raw + (int64_t)(i * J + j) * 3;
            r[0] = X[0]; r[1] = X[1]; r[2] = X[2];
        }
    }
    PAM_FOR_REV(cam, V) {
        int k = 0;
        const int mm = sh.m[cam];
        PAM_NOUNROLL for (int d = 0; d < mm; ++d)
            if (sh.um_flag[cam][d]) sh.um[cam][k++] = (signed char)d;
        sh.um_n[cam] = (signed char)k;
    }
    ctx.phase_sync();
    // new-track initialisation has something to do when two cameras hold unmatched detections (a hypothesis
    // needs views from two cameras to become a track, hypothesis.size() > 1); every thread evaluates it
    // (uniform), so the staged detections can be released right here on the frames that do not need them
    int do_init;
    {
        int cams_with = 0;
        PAM_NOUNROLL for (int cam = 0; cam < V; ++cam) cams_with += (sh.um_n[cam] > 0) ? 1 : 0;
        do_init = (V >= 2 && cams_with >= 2) ? 1 : 0;
    }
    if (!do_init) hook.dets_released();

    // ---- phase 6: per (track, joint) of every successfully updated track: temporal Gaussian, last
    //      sample (IterativeTracker.py:371-383), history append, velocity = float32 mean of the
    //      last <= 5 differences (:385-395), output row (ivclabpose.py:265-287) ------------------
    {
        double* const rawb = sq.raw_;
        PAM_FOR(it, n * J) {
            const int i = fast_div(it, c.inv_J), j = it - i * J;
            if (!track_ok(c, sq, i)) continue;
            const int s = hdr.order[i];
            TrkMeta& t = trk[s];
            const int which = (c.arm_mask >> j) & 1u;
            const int rad = c.rad[which];
            const double* w = c.gw[which];
            const int L = t.hist_len, N = L + 1;       // series = history + current raw pose
            const int start = t.hist_start;
            const int pos = (start + L) % PAM_HIST;
            double* raw = rawb + (int64_t)(i * J + j) * 3;
            const double* hb = g.hist + ((int64_t)(s * PAM_HIST) * J + j) * 3;   // + ring * J3
            const double raw0 = raw[0], raw1 = raw[1], raw2 = raw[2];
            double o0 = raw0 * w[0], o1 = raw1 * w[0], o2 = raw2 * w[0];
            // window after the append and the at-most-one-entry trim (IterativeTracker.py:330-332)
            int len2 = L + 1, start2 = start;
            if (frame - t.hist_time[start] > c.max_age) { start2 = (start + 1) % PAM_HIST; len2 -= 1; }
            float* vel = g.vel + (int64_t)(s * J + j) * 3;
            if (rad <= PAM_RECENT && rad <= L) {
                // Steady state.  Both the Gaussian (samples L-rad .. L, each history sample used twice by the
                // reflection at the end of the series) and the velocity (last <= 5 differences) read the newest
                // history entries: they are fetched once, all loads in flight together -- the ring was written
                // by this group's own global stores, so every dependent load would pay an L2 round trip.
                double hx[PAM_RECENT][3];
#pragma unroll
                for (int q = 0; q < PAM_RECENT; ++q) {
                    const int e = (q < L) ? L - 1 - q : 0;                       // clamped: never used beyond L
                    const double* x = hb + (int64_t)((start + e) % PAM_HIST) * J3;
                    hx[q][0] = x[0]; hx[q][1] = x[1]; hx[q][2] = x[2];
                }
#pragma unroll
                for (int k = PAM_RECENT; k >= 1; --k) {                          // same order as the generic loop
                    if (k > rad) continue;
                    // sample L-k and its mirror partner L-k+1 (the current raw pose for k = 1)
                    const double r0 = (k == 1) ? raw0 : hx[k >= 2 ? k - 2 : 0][0];
                    const double r1 = (k == 1) ? raw1 : hx[k >= 2 ? k - 2 : 0][1];
                    const double r2 = (k == 1) ? raw2 : hx[k >= 2 ? k - 2 : 0][2];
                    o0 += (hx[k - 1][0] + r0) * w[k];
                    o1 += (hx[k - 1][1] + r1) * w[k];
                    o2 += (hx[k - 1][2] + r2) * w[k];
                }
                if (len2 >= 2) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
                    float h0 = (float)o0, h1 = (float)o1, h2 = (float)o2;    // newest entry = this frame's pose
                    const int cnt = (len2 - 1 < 5) ? len2 - 1 : 5;
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        if (q >= cnt) continue;
                        const float l0 = (float)hx[q][0], l1 = (float)hx[q][1], l2 = (float)hx[q][2];
                        a0 += h0 - l0; a1 += h1 - l1; a2 += h2 - l2;
                        h0 = l0; h1 = l1; h2 = l2;
                    }
                    const float fc = (float)cnt;
                    vel[0] = a0 / fc; vel[1] = a1 / fc; vel[2] = a2 / fc;
                }
            } else {
                PAM_NOUNROLL for (int k = rad; k >= 1; --k) {
                    const int il = reflect_index(L - k, N), ir = reflect_index(L + k, N);
                    const double* xl = (il == L) ? raw : hb + (int64_t)((start + il) % PAM_HIST) * J3;
                    const double* xr = (ir == L) ? raw : hb + (int64_t)((start + ir) % PAM_HIST) * J3;
                    o0 += (xl[0] + xr[0]) * w[k];
                    o1 += (xl[1] + xr[1]) * w[k];
                    o2 += (xl[2] + xr[2]) * w[k];
                }
                if (len2 >= 2) {
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
                    float h0 = (float)o0, h1 = (float)o1, h2 = (float)o2;    // newest entry = this frame's pose
                    int cnt = 0;
                    PAM_NOUNROLL for (int idx = len2 - 1; idx >= 1 && cnt < 5; --idx, ++cnt) {
                        const double* lo = hb + (int64_t)((start2 + idx - 1) % PAM_HIST) * J3;
                        const float l0 = (float)lo[0], l1 = (float)lo[1], l2 = (float)lo[2];
                        a0 += h0 - l0; a1 += h1 - l1; a2 += h2 - l2;
                        h0 = l0; h1 = l1; h2 = l2;
                    }
                    const float fc = (float)cnt;
                    vel[0] = a0 / fc; vel[1] = a1 / fc; vel[2] = a2 / fc;
                }
            }
            double* dst = g.hist + ((int64_t)(s * PAM_HIST + pos) * J + j) * 3;
            dst[0] = o0; dst[1] = o1; dst[2] = o2;
            g.nv[s * J + j] = sh.nvj[i][j];
            if (j == 0) t.hist_time[pos] = frame;     // no other thread reads this entry in this phase
            raw[0] = o0; raw[1] = o1; raw[2] = o2;    // smoothed joint, for the output rows written below
        }
    }
    // ... while one otherwise idle thread ranks the tracks that are reported this frame (output row of each)
    if (ctx.tid() == ctx.nthreads() - 1) {
        int k = 0;
        PAM_NOUNROLL for (int i = 0; i < n; ++i) sh.out_row[i] = track_reported(c, sq, i) ? (signed char)(k++) : (signed char)-1;
        if (out.count) *out.count = k;
    }
    ctx.phase_sync();
    {
        const double* const rawb = sq.raw_;
        PAM_FOR(it, n * J) {
            const int i = fast_div(it, c.inv_J), j = it - i * J;
            const int k = sh.out_row[i];
            if (k < 0 || k >= c.max_rep) continue;       // rows beyond the output stride are dropped (count tells)
            const double* o = rawb + (int64_t)(i * J + j) * 3;
            if (out.joints) {
                float* oj = out.joints + (int64_t)(k * J + j) * 3;
                oj[0] = (float)o[0]; oj[1] = (float)o[1]; oj[2] = (float)o[2];
            }
            if (out.nviews) out.nviews[k * J + j] = sh.nvj[i][j];
        }
    }

    // ---- phase 7: life-cycle (IterativeTracker.py:253-274), reported ids, reap (:178) -----------
    // 7a, one thread per track: counters and state transitions; bit 0 of life_flag = keep, bit 1 = reported
    PAM_FOR_REV(i, n) {
        const int s = hdr.order[i];
        TrkMeta& t = trk[s];
        if (sh.do_update[i] && !((double)sh.fail[i] > c.fail_limit) && t.hist_len >= PAM_HIST)
            hdr.status = SEQ_ERR_HIST_OVERFLOW;
        int flag = 0;
        if (track_ok(c, sq, i)) {
            t.hist_len += 1;
            if (frame - t.hist_time[t.hist_start] > c.max_age) {
                t.hist_start = (t.hist_start + 1) % PAM_HIST;
                t.hist_len -= 1;
            }
            t.hits += 1;
            t.tsu = 0;
            t.vt_last = sh.gv_n[i];
            if (t.state == ST_TENTATIVE && t.hits >= c.n_init) t.state = ST_CONFIRMED;
            if (t.state == ST_CONFIRMED) flag |= 2;
        } else {
            if (t.state == ST_TENTATIVE && !t.already) t.state = ST_DELETED;
            else if (t.tsu >= c.max_age) t.state = ST_DELETED;
        }
        if (t.state != ST_DELETED) flag |= 1; else fs.any_deleted = 1;
        if ((flag & 2) && out.ids && sh.out_row[i] < c.max_rep) out.ids[sh.out_row[i]] = t.track_id;     // reported <=> out_row >= 0
        if ((flag & 2) && out.vlist && sh.out_row[i] < c.max_rep) {
            unsigned char* vl = out.vlist + sh.out_row[i] * PAM_VLIST;
            vl[0] = (unsigned char)sh.gv_n[i];
            vl[1] = (unsigned char)t.nviews;
            PAM_NOUNROLL for (int k = 0; k < PAM_MAX_V; ++k)       // the whole row is written (unused entries 0)
                vl[2 + k] = (k < t.nviews) ? (unsigned char)(t.view_cid[k] | (t.view_time[k] == frame ? 0x80 : 0)) : (unsigned char)0;
        }
        sh.life_flag[i] = (signed char)flag;
    }
    // ... while the thread below them seeds the hypothesis list of the initialisation (if any)
    if (ctx.tid() == (ctx.nthreads() - 1 - n > 0 ? ctx.nthreads() - 1 - n : 0)) {
        fs.hyp_n = 0;
        if (do_init) {
            const int n0 = sh.um_n[0];
            PAM_NOUNROLL for (int q = 0; q < n0; ++q) {
                sh.hyp_nviews[q] = 1; sh.hyp_cam[q][0] = 0; sh.hyp_det[q][0] = sh.um[0][q];
            }
            fs.hyp_n = n0;
        }
    }
    ctx.phase_sync();
    // 7b, one thread: compaction of the track list, only on the frames a track was deleted
    if (ctx.tid() == 0) {
        if (fs.any_deleted) {
            int wr = 0;
            PAM_NOUNROLL for (int i = 0; i < n; ++i) {
                const int s = hdr.order[i];
                if (sh.life_flag[i] & 1) hdr.order[wr++] = (signed char)s;
                else hdr.used_mask &= ~(1u << s);
            }
            hdr.ntracks = wr;
        }
        hdr.frames_done += 1;
    }
    if (out.timing) tk2 = ctx.clock();

    // ---- phase 8: new-track initialisation (IterativeTracker.py:52-113); rare in steady state ---
    if (do_init) {             // uniform
        ctx.sync();            // the track list of 7b is read (and extended) below
        double* const hcost = sq.hyp_cost(c);
        unsigned char* const hveto = sq.hyp_veto(c);
        // grow hypotheses camera by camera
        PAM_NOUNROLL for (int cam = 1; cam < V; ++cam) {
            const int nh = fs.hyp_n, nd = sh.um_n[cam];
            if (nd == 0) continue;           // uniform
            const signed char* um = sh.um[cam];
            PAM_FOR(it, nh * nd) {
                const int h = it / nd, p = it % nd;
                bool veto;
                hcost[h * D + p] = hyp_cost(c, sq, dets, h, cam, um[p], veto);
                hveto[h * D + p] = veto ? 1 : 0;
            }
            ctx.sync();
            if (ctx.tid() == 0) {
                int col4row[PAM_LSAP_N];
                uint32_t handled = 0u;
                lsap_solve<PAM_LSAP_N>(nh, nd, [&](int h, int p) { return hcost[h * D + p]; }, col4row);
                int hn = nh;
                auto spawn = [&](int p) {          // a new single-view hypothesis from detection um[p]
                    if (hn >= c.max_hyp) { hdr.warn |= WARN_HYP_OVERFLOW; fs.warned = 1; return; }
                    sh.hyp_nviews[hn] = 1; sh.hyp_cam[hn][0] = (signed char)cam; sh.hyp_det[hn][0] = um[p];
                    ++hn;
                };
                PAM_NOUNROLL for (int h = 0; h < nh; ++h) {
                    const int p = col4row[h];
                    if (p < 0) continue;
                    handled |= (1u << p);
                    if (hveto[h * D + p]) spawn(p);
                    else {
                        const int k = sh.hyp_nviews[h]++;
                        sh.hyp_cam[h][k] = (signed char)cam; sh.hyp_det[h][k] = um[p];
                    }
                }
                PAM_NOUNROLL for (int p = 0; p < nd; ++p)
                    if (!((handled >> p) & 1u)) spawn(p);
                fs.hyp_n = hn;
            }
            ctx.sync();
        }
        // first triangulation of every multi-view hypothesis (hypothesis.py:23-44)
        const int nh = fs.hyp_n;
        PAM_FOR(h, nh) sh.hyp_fail[h] = (sh.hyp_nviews[h] < 2) ? 1 : 0;
        ctx.sync();
        PAM_FOR(it, nh * J) {
            const int h = it / J, j = it % J;
            const int Vt = sh.hyp_nviews[h];
            if (Vt < 2) continue;
            double u[PAM_MAX_V], v[PAM_MAX_V];
            const signed char* hcam = sh.hyp_cam[h];
            const signed char* hdet = sh.hyp_det[h];
            PAM_NOUNROLL for (int a = 0; a < Vt; ++a) {
                const float* q = dets + ((int64_t)(hcam[a] * D + hdet[a]) * J + j) * 3;
                v[a] = (double)q[0];
                u[a] = (double)q[1];
            }
            double X[3] = {0.0, 0.0, 0.0};
            const int nv = joint_init(c, sq, Vt, hcam, u, v, X);
            sq.hyp_nvj(c)[h * J + j] = (unsigned char)nv;
            if (nv < 2) sh.hyp_fail[h] = 1;     // benign race: every writer stores 1
            double* r = sq.hyp_pose(c) + (int64_t)(h * J + j) * 3;
            r[0] = X[0]; r[1] = X[1]; r[2] = X[2];
        }
        ctx.sync();
        // spawn tracks in hypothesis order (IterativeTracker.py:102-113)
        if (ctx.tid() == 0) {
            PAM_NOUNROLL for (int h = 0; h < nh; ++h) {
                sh.hyp_slot[h] = -1;
                if (sh.hyp_fail[h] || hdr.status != SEQ_OK) continue;
                int s = 0;
                PAM_NOUNROLL while (s < MT && ((hdr.used_mask >> s) & 1u)) ++s;
                if (s >= MT || hdr.ntracks >= MT) { hdr.warn |= WARN_TRACK_OVERFLOW; fs.warned = 1; continue; }
                hdr.used_mask |= (1u << s);
                hdr.order[hdr.ntracks++] = (signed char)s;
                TrkMeta& t = trk[s];
                t.track_id = hdr.next_id++;
                t.hits = 1; t.age = 1; t.tsu = 0; t.state = ST_TENTATIVE; t.already = 0;
                t.nviews = sh.hyp_nviews[h];
                t.vt_last = t.nviews;
                PAM_NOUNROLL for (int cc2 = 0; cc2 < PAM_MAX_V; ++cc2) t.view_slot[cc2] = -1;
                PAM_NOUNROLL for (int k = 0; k < t.nviews; ++k) {
                    const int hc = sh.hyp_cam[h][k];
                    t.view_cid[k] = hc; t.view_time[k] = frame; t.view_slot[hc] = (signed char)k;
                    t.view_det[k] = sh.hyp_det[h][k];
                }
                t.hist_start = 0; t.hist_len = 1; t.hist_time[0] = frame;
                sh.hyp_slot[h] = (signed char)s;
            }
        }
        ctx.sync();
        PAM_FOR(it, nh * J3) {
            const int h = it / J3, e = it % J3;
            const int s = sh.hyp_slot[h];
            if (s < 0) continue;
            g.hist[(int64_t)(s * PAM_HIST) * J3 + e] = sq.hyp_pose(c)[(int64_t)h * J3 + e];
            g.vel[(int64_t)s * J3 + e] = 0.0f;
            if (e < J) g.nv[s * J + e] = sq.hyp_nvj(c)[h * J + e];
        }
        hook.dets_released();
    }
    if (ctx.tid() == 0) {
        if (fs.warned) { hdr.warn_frames += 1; fs.warned = 0; }   // ordered after its last writer by the syncs above
        if (out.timing) {
            const long long tk3 = ctx.clock();
            out.timing[0] = (int)(tk1 - tk0); out.timing[1] = (int)(tk2 - tk1); out.timing[2] = (int)(tk3 - tk2);
            out.timing[3] = (int)(tk3 - tk0);
        }
    }
}

// End of a launch: write the (v, u, conf) triples of every view that was matched during this launch
// into the track state, so that later launches (and the host-side state read-back) find them there.
template <class Ctx, class K>
PAM_HD void persist_views(Ctx& ctx, const DevCfg& c, const Seq<K>& sq, const float* gin, int gin_frame0,
                          const float* last_staged, int last_tl) {
    const int V = c.V, D = c.D, J3 = c.J * 3;
    const int lanes = (ctx.nthreads() >= 32) ? 32 : ctx.nthreads();
    const int grp = ctx.tid() / lanes, ngrp = ctx.nthreads() / lanes, lane = ctx.tid() - grp * lanes;
    const SeqShared<K>& sh = *sq.sh;
    const SeqHeader& hdr = sh.hdr;
    const int n = hdr.ntracks;
    PAM_NOUNROLL for (int p = grp; p < n * V; p += ngrp) {
        const int i = fast_div(p, c.inv_V), k = p - i * V;
        const int s = hdr.order[i];
        const TrkMeta& t = sh.trk[s];
        if (k >= t.nviews) continue;
        const int tl = t.view_time[k] - gin_frame0;
        if (tl < 0) continue;                                   // matched in an earlier launch: already stored
        // the launch's last frame may still be staged on chip (for one-frame launches the input may even live in
        // mapped host memory)
        const float* src = (tl == last_tl && last_staged) ? last_staged + (int64_t)(t.view_cid[k] * D + t.view_det[k]) * J3
                                                          : gin + ((int64_t)tl * V * D + t.view_cid[k] * D + t.view_det[k]) * J3;
        float* dst = sq.g.view + (int64_t)(s * V + k) * J3;
        PAM_NOUNROLL for (int e = lane; e < J3; e += lanes) dst[e] = src[e];
    }
}

}  // namespace pam
