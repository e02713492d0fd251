// pam_lib.cu -- sm_100a kernels + the C ABI of libpam.so (include/pam.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//        (see __graft_entry__.build()).  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pam_host.h"
#include "pam_ops.cuh"

using namespace pam;

// ----------------------------------------------------------------------------------------------
// persistent per-sequence tracker kernel: one CTA owns one sequence for all T frames
// ----------------------------------------------------------------------------------------------
template <int AFF_UNROLL>
struct DeviceCtxT {
    static constexpr int kAffinityUnroll = AFF_UNROLL;
    __host__ __device__ __forceinline__ int tid() const {
#ifdef __CUDA_ARCH__
        return threadIdx.x;
#else
        return 0;
#endif
    }
    __host__ __device__ __forceinline__ int nthreads() const {
#ifdef __CUDA_ARCH__
        return blockDim.x;
#else
        return 1;
#endif
    }
    __host__ __device__ __forceinline__ void sync() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    __host__ __device__ __forceinline__ void atomic_inc(int* p) const {
#ifdef __CUDA_ARCH__
        atomicAdd(p, 1);
#else
        *p += 1;
#endif
    }
};

// Asynchronous global -> shared staging of one frame's detections, issued for frame t+1 while frame t is
// being processed, so the frame-serial chain never waits on HBM.  When the frame is a whole number of
// 16-byte units (Shelf: 3360 B) ONE thread issues ONE bulk copy through the TMA engine
// (cp.async.bulk, completion counted in bytes on an mbarrier -- UBLKCP in SASS); otherwise all threads
// issue 4-byte LDGSTS copies.  The V per-camera counts always go through LDGSTS.
__device__ __forceinline__ bool stage_is_bulk(const float* gsrc, int nfloats) {
    return (nfloats & 3) == 0 && ((uintptr_t)gsrc & 15) == 0;
}
__device__ __forceinline__ void stage_frame(float* sdst, int* scnt, const float* gsrc, const int* gcnt, int nfloats,
                                            int V, unsigned long long* mbar, bool bulk) {
    if (bulk) {
        if (threadIdx.x == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sdst), bar = (unsigned)__cvta_generic_to_shared(mbar);
            const unsigned bytes = (unsigned)nfloats * 4u;
            // order the generic-proxy reads of this buffer (two frames ago) before the async-proxy write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
        }
    } else {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sdst);
        PAM_NOUNROLL for (int i = threadIdx.x; i < nfloats; i += blockDim.x)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + i * 4), "l"(gsrc + i) : "memory");
    }
    const unsigned cbase = (unsigned)__cvta_generic_to_shared(scnt);
    PAM_NOUNROLL for (int i = threadIdx.x; i < V; i += blockDim.x)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cbase + i * 4), "l"(gcnt + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// every thread: the LDGSTS copies of this thread are done; the bulk copy (if any) has delivered all bytes
__device__ __forceinline__ void stage_wait(unsigned long long* mbar, unsigned parity, bool bulk) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (bulk) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(mbar);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "PAM_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@!p bra PAM_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
    }
}

// dynamic shared memory of k_track_sequences: [arena doubles][2 x frame floats][2 x V counts]
static inline size_t frame_floats_padded(const DevCfg& c) { return ((size_t)c.V * c.D * c.J * 3 + 3) / 4 * 4; }
static inline size_t track_smem_bytes(const DevCfg& c) {
    return (size_t)((arena_doubles(c) + 1) / 2 * 2) * 8 + 2 * frame_floats_padded(c) * 4 + 2 * PAM_MAX_V * 4;
}

struct TrackIO {
    const float* dets;      // [S][T][V][D][J][3]
    const int* counts;      // [S][T][V]
    int* out_count;         // [S][T]
    int* out_ids;           // [S][T][MT]
    float* out_joints;      // [S][T][MT][J][3]
    unsigned char* out_nv;  // [S][T][MT][J]
    int* out_assoc;         // [S][T][V][D]
    int seq_frames;         // frames between consecutive sequences in every tensor above (>= T)
    int* out_status;        // [S] final status word of each sequence, or null
};

#define PAM_TRACK_THREADS_MAX 256

template <int MAXT, int MINB, int TEAM>
__global__ void __launch_bounds__(MAXT, MINB)
k_track_sequences(const DevCfg c, const CamConst cc, char* __restrict__ state, int T, int frame0, const TrackIO io) {
    extern __shared__ double arena[];
    __shared__ SeqShared sh;
    DeviceCtxT<(MAXT * MINB <= 512) ? 4 : 2> ctx;     // <= 512 resident threads per SM: 128 registers each
    const int s = blockIdx.x;
    SeqGlobal g;
    g.bind(c, state + (int64_t)s * c.seq_bytes);
    const int nfl = c.V * c.D * c.J * 3;
    const int nfl_pad = (nfl + 3) / 4 * 4;
    float* dbuf = (float*)(arena + (arena_doubles(c) + 1) / 2 * 2);   // 16-byte aligned for cp.async
    int* cbuf = (int*)(dbuf + 2 * nfl_pad);
    const float* gd = io.dets + (int64_t)s * io.seq_frames * nfl;
    const int* gc = io.counts + (int64_t)s * io.seq_frames * c.V;
    __shared__ __align__(8) unsigned long long mbar[2];       // one transaction barrier per detection buffer
    const bool bulk = stage_is_bulk(gd, nfl);
    if (threadIdx.x == 0) {
        carve(c, sh, arena, g);
        const unsigned b0 = (unsigned)__cvta_generic_to_shared(&mbar[0]), b1 = (unsigned)__cvta_generic_to_shared(&mbar[1]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (T > 0) stage_frame(dbuf, cbuf, gd, gc, nfl, c.V, &mbar[0], bulk);
    load_cameras(ctx, c, sh, cc);
    load_state(ctx, c, sh, g);
    if (T > 0) stage_wait(&mbar[0], 0u, bulk);
    __syncthreads();
#if defined(PAM_PHASE_TIMING)
    if (threadIdx.x == 0) { for (int k = 0; k < 24; ++k) sh.phase_cyc[k] = 0; sh.tlast = clock64(); }
#endif
    FrameOut o;
    {
        const int64_t f0 = (int64_t)s * io.seq_frames;
        o.count = io.out_count ? io.out_count + f0 : nullptr;
        o.ids = io.out_ids ? io.out_ids + f0 * c.max_trk : nullptr;
        o.joints = io.out_joints ? io.out_joints + f0 * c.max_trk * c.J * 3 : nullptr;
        o.nviews = io.out_nv ? io.out_nv + f0 * c.max_trk * c.J : nullptr;
        o.assoc = io.out_assoc ? io.out_assoc + f0 * c.V * c.D : nullptr;
    }
    const int st_ids = c.max_trk, st_joints = c.max_trk * c.J * 3, st_nv = c.max_trk * c.J, st_assoc = c.V * c.D;
    PAM_NOUNROLL for (int t = 0; t < T; ++t) {
        const int cur = t & 1;
        if (t + 1 < T)
            stage_frame(dbuf + (cur ^ 1) * nfl_pad, cbuf + (cur ^ 1) * PAM_MAX_V, gd + (int64_t)(t + 1) * nfl,
                        gc + (t + 1) * c.V, nfl, c.V, &mbar[cur ^ 1], bulk);
        if (TEAM > 1) frame_step<WarpTeam<TEAM>>(ctx, c, sh, g, frame0 + t, dbuf + cur * nfl_pad, cbuf + cur * PAM_MAX_V, o, gd, frame0);
        else frame_step<SoloTeam>(ctx, c, sh, g, frame0 + t, dbuf + cur * nfl_pad, cbuf + cur * PAM_MAX_V, o, gd, frame0);
        if (o.count) o.count += 1;
        if (o.ids) o.ids += st_ids;
        if (o.joints) o.joints += st_joints;
        if (o.nviews) o.nviews += st_nv;
        if (o.assoc) o.assoc += st_assoc;
        // buffer (t+1)&1 is used for the ((t+1)>>1)-th time: that is the parity of its barrier phase
        if (t + 1 < T) stage_wait(&mbar[cur ^ 1], (unsigned)(((t + 1) >> 1) & 1), bulk);
        __syncthreads();
        PAM_MARK(8);
    }
    persist_views(ctx, c, sh, g, gd, frame0, T > 0 ? dbuf + ((T - 1) & 1) * nfl_pad : nullptr, T - 1);
    store_state(ctx, c, sh, g);
    if (io.out_status && threadIdx.x == 0) {
        // the status word may live in mapped host memory and be polled by the caller (small-job path):
        // everything this CTA wrote for the host (ordered before this point by the frame barriers) must be
        // visible system-wide first
        __threadfence_system();
        *(volatile int*)(io.out_status + s) = sh.hdr.status;
    }
#if defined(PAM_PHASE_TIMING)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        static const char* nm[9] = {"1 age+reproj", "2 affinity", "3 assign", "4 add_pose+believe", "5 filter+dlt",
                                    "6 smooth+motion+out", "7 lifecycle+reap", "8 init", "9 stage wait"};
        long long tot = 0;
        for (int k = 0; k < 9; ++k) tot += sh.phase_cyc[k];
        for (int k = 0; k < 9; ++k)
            printf("phase %-22s %9.0f cyc/frame %5.1f%%\n", nm[k], (double)sh.phase_cyc[k] / T, 100.0 * sh.phase_cyc[k] / tot);
        static const char* sub[7] = {"5a pair tests", "5b conflict resolution", "5c gram fold", "5d solve",
                                     "6a prologue", "6b smoothing", "6c velocity"};
        for (int k = 0; k < 7; ++k)   // thread 0's item; the remainder of phase 5 (stores, barrier wait) stays in its row above
            printf("  sub %-22s %9.0f cyc/frame\n", sub[k], (double)sh.phase_cyc[10 + k] / T);
        printf("total %.0f cyc/frame\n", (double)tot / T);
    }
#endif
}

// ----------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct pam_handle {
    pam_config cfg;
    DevCfg dc;
    int device = 0;
    bool have_cameras = false;
    int track_threads = 128;
    bool threads_forced = false;
    int track_minblocks = 0;   // 0 = choose per launch
    int num_sms = 148;
    int track_team = 0;        // 0 = choose per launch
    DevBuf cam;              // packed camera constants
    CamConst cc{};
    // workspace of the *_host entry points
    DevBuf ws_state, ws_dets, ws_counts, ws_count, ws_ids, ws_joints, ws_nv, ws_assoc;
    int ws_S = 0;
    bool smem_opt_in = false;
    cudaStream_t ws_stream = nullptr, ws_in = nullptr, ws_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    char* zc = nullptr;        // pinned, device-mapped staging of the small-job path
    size_t zc_cap = 0;
    int64_t launches = 0;
    std::string err;
};

static thread_local std::string g_err;

static int fail(pam_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    g_err = msg;
    return code;
}
static int cuda_fail(pam_handle* h, cudaError_t e, const char* what) {
    return fail(h, PAM_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return cuda_fail(h, _e, #call);     \
    } while (0)

extern "C" {

int pam_abi_version(void) { return PAM_ABI_VERSION; }

const char* pam_status_string(int s) {
    switch (s) {
        case PAM_OK: return "ok";
        case PAM_E_INVALID: return "invalid argument or configuration";
        case PAM_E_CUDA: return "CUDA runtime error";
        case PAM_E_NOCAMERAS: return "pam_set_cameras has not been called";
        case PAM_E_CAPACITY: return "a sequence exceeded a capacity limit (tracks / hypotheses / detections)";
        case PAM_E_INTERNAL: return "internal error";
        default: return "unknown status";
    }
}

const char* pam_last_error(const pam_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int pam_create(const pam_config* cfg, int device, pam_handle** out) {
    if (!cfg || !out) return fail(nullptr, PAM_E_INVALID, "null argument");
    *out = nullptr;
    pam_handle* h = new (std::nothrow) pam_handle();
    if (!h) return fail(nullptr, PAM_E_INTERNAL, "out of host memory");
    h->cfg = *cfg;
    h->device = device;
    std::string err;
    int rc = make_devcfg(*cfg, h->dc, err);
    if (rc != PAM_OK) { delete h; return fail(nullptr, rc, err); }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        delete h;
        return fail(nullptr, PAM_E_CUDA, std::string("no usable CUDA device (libpam has no CPU fallback): ") +
                                             (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaSetDevice"); }
    const char* nt = getenv("PAM_TRACK_THREADS");
    if (nt) {
        int v = atoi(nt);
        if (v >= 32 && v <= PAM_TRACK_THREADS_MAX) { h->track_threads = (v / 32) * 32; h->threads_forced = true; }
    } else {
        int want = cfg->max_tracks * cfg->num_joints;     // one thread per (track, joint)
        h->track_threads = want <= 64 ? 64 : (want <= 128 ? 128 : (want <= 192 ? 192 : 256));
    }
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    const char* tm = getenv("PAM_TRACK_TEAM");
    if (tm) { int v = atoi(tm); if (v == 1 || v == 2) h->track_team = v; }
    const char* mb = getenv("PAM_TRACK_MINBLOCKS");
    if (mb) { int v = atoi(mb); if (v == 4 || v == 6 || v == 8 || v == 10 || v == 12) h->track_minblocks = v; }
    *out = h;
    return PAM_OK;
}

int pam_destroy(pam_handle* h) {
    if (!h) return PAM_OK;
    cudaSetDevice(h->device);
    h->cam.release();
    h->ws_state.release(); h->ws_dets.release(); h->ws_counts.release(); h->ws_count.release();
    h->ws_ids.release(); h->ws_joints.release(); h->ws_nv.release(); h->ws_assoc.release();
    if (h->ws_stream) cudaStreamDestroy(h->ws_stream);
    if (h->ws_in) cudaStreamDestroy(h->ws_in);
    if (h->ws_out) cudaStreamDestroy(h->ws_out);
    for (auto e : h->ev_in) cudaEventDestroy(e);
    for (auto e : h->ev_k) cudaEventDestroy(e);
    if (h->zc) cudaFreeHost(h->zc);
    delete h;
    return PAM_OK;
}

int pam_set_cameras(pam_handle* h, const float* P, const float* RKinv, const double* pos, const float* F) {
    if (!h || !P || !RKinv || !pos || !F) return fail(h, PAM_E_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    const int V = h->cfg.num_cameras;
    const size_t bpos = (size_t)V * 3 * 8, bP = (size_t)V * 12 * 4, bRK = (size_t)V * 9 * 4, bF = (size_t)V * V * 9 * 4;
    CK(h->cam.reserve(bpos + bP + bRK + bF));
    char* base = (char*)h->cam.p;
    CK(cudaMemcpy(base, pos, bpos, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(base + bpos, P, bP, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(base + bpos + bP, RKinv, bRK, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(base + bpos + bP + bRK, F, bF, cudaMemcpyHostToDevice));
    h->cc.pos = (const double*)base;
    h->cc.P = (const float*)(base + bpos);
    h->cc.RKinv = (const float*)(base + bpos + bP);
    h->cc.F = (const float*)(base + bpos + bP + bRK);
    h->have_cameras = true;
    return PAM_OK;
}

int pam_get_state_layout(const pam_handle* h, pam_state_layout* out) {
    if (!h || !out) return PAM_E_INVALID;
    fill_layout(h->dc, *out);
    return PAM_OK;
}

int pam_track_reset(pam_handle* h, void* d_state, int32_t S, void* stream) {
    if (!h || !d_state || S < 0) return fail(h, PAM_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(d_state, 0, (size_t)h->dc.seq_bytes * S, (cudaStream_t)stream));
    return PAM_OK;
}

typedef void (*track_kernel_t)(const DevCfg, const CamConst, char*, int, int, const TrackIO);

// register budget variants: <= 128 threads with 4 / 6 / 8 CTAs per SM, or up to 256 threads
// Launch shape per call.  Few sequences: latency matters -- one CTA per sequence with the full register
// budget.  Many sequences: the kernel is latency/barrier bound, so more, smaller CTAs per SM win
// (measured on B200, Shelf shape, final kernel: 128 thr x 8/SM 57 M frames/s, 96 x 8 59 M, 64 x 12 59 M at 12 per SM).
// Two lanes per (track, joint) (PAM_TRACK_TEAM=2) measured no faster than one, so 1 is the default.
static track_kernel_t pick_track_kernel(const pam_handle* h, int S, int* threads) {
    const int per_sm = (S + h->num_sms - 1) / h->num_sms;
    const int team = h->track_team ? h->track_team : 1;
    int nt = h->track_threads;                       // PAM_TRACK_THREADS or the size-based default
    int mb = h->track_minblocks;
    if (!h->threads_forced && nt <= 128 && team == 1 && !mb && per_sm >= 10) nt = 64;
    // 7-9 sequences per SM: three warps per CTA (the fourth one of a 128-thread CTA has nothing to do in any
    // phase of the usual shapes) leave 80 registers per thread at 8 CTAs per SM: 59 M against 57 M frames/s
    if (!h->threads_forced && nt == 128 && team == 1 && !mb && per_sm >= 7 && per_sm < 10) nt = 96;
    if (!mb) mb = nt > 128 ? (per_sm > 2 ? 4 : 2) : (nt <= 64 && per_sm >= 10 ? 12 : (per_sm >= 7 ? 8 : (per_sm > 4 ? 6 : 4)));
    *threads = nt;
    if (nt > 128) {
        if (team > 1) return k_track_sequences<256, 2, 2>;
        return mb >= 4 ? k_track_sequences<256, 4, 1> : k_track_sequences<256, 2, 1>;
    }
    if (team > 1) return mb >= 6 ? k_track_sequences<128, 6, 2> : k_track_sequences<128, 4, 2>;
    if (nt == 96 && team == 1) return k_track_sequences<96, 8, 1>;     // three warps, 85 registers, 8 CTAs per SM
    if (nt <= 64 && mb >= 10) return mb >= 12 ? k_track_sequences<64, 12, 1> : k_track_sequences<64, 10, 1>;
    switch (mb) {
        case 8: case 10: case 12: return k_track_sequences<128, 8, 1>;
        case 6: return k_track_sequences<128, 6, 1>;
        default: return k_track_sequences<128, 4, 1>;
    }
}

static int launch_track(pam_handle* h, void* d_state, int32_t S, int32_t T, int32_t frame0, const TrackIO& io,
                        cudaStream_t stream) {
    const size_t smem = track_smem_bytes(h->dc);
    int threads = h->track_threads;
    track_kernel_t kern = pick_track_kernel(h, S, &threads);
    if (smem + sizeof(SeqShared) > 48 * 1024)
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<S, threads, smem, stream>>>(h->dc, h->cc, (char*)d_state, T, frame0, io);
    h->launches += 1;
    CK(cudaGetLastError());
    return PAM_OK;
}

int pam_track_sequences(pam_handle* h, void* d_state, int32_t S, int32_t T, int32_t frame0, const float* d_dets,
                        const int32_t* d_counts, int32_t* d_out_count, int32_t* d_out_ids, float* d_out_joints,
                        uint8_t* d_out_nviews, int32_t* d_out_assoc, void* stream) {
    if (!h || !d_state || !d_dets || !d_counts || S < 0 || T < 0) return fail(h, PAM_E_INVALID, "bad argument");
    if (!h->have_cameras) return fail(h, PAM_E_NOCAMERAS, "pam_set_cameras has not been called");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    if (S == 0 || T == 0) return PAM_OK;
    CK(cudaSetDevice(h->device));
    TrackIO io{d_dets, d_counts, d_out_count, d_out_ids, d_out_joints, d_out_nviews, d_out_assoc, T, nullptr};
    return launch_track(h, d_state, S, T, frame0, io, (cudaStream_t)stream);
}

int pam_track_status(pam_handle* h, const void* d_state, int32_t S, int32_t* h_status, void* stream) {
    if (!h || !d_state || S < 0) return fail(h, PAM_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    std::vector<int32_t> tmp((size_t)S);
    const char* base = (const char*)d_state + h->dc.off_hdr + offsetof(SeqHeader, status);
    CK(cudaMemcpy2DAsync(tmp.data(), 4, base, (size_t)h->dc.seq_bytes, 4, (size_t)S, cudaMemcpyDeviceToHost,
                         (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    int bad = 0, first = -1, code = 0;
    for (int s = 0; s < S; ++s) {
        if (h_status) h_status[s] = tmp[s];
        if (tmp[s] != 0) { if (!bad) { first = s; code = tmp[s]; } ++bad; }
    }
    if (bad) {
        static const char* names[] = {"ok", "track slots exhausted (raise max_tracks)", "hypothesis table exhausted",
                                      "more detections than max_detections", "pose history ring exhausted"};
        char msg[256];
        snprintf(msg, sizeof msg, "%d sequence(s) hit a capacity limit; first: sequence %d: %s", bad, first,
                 (code >= 0 && code <= 4) ? names[code] : "unknown");
        return fail(h, PAM_E_CAPACITY, msg);
    }
    return PAM_OK;
}

int pam_track_sequences_host(pam_handle* h, int32_t S, int32_t T, int32_t frame0, int32_t fresh, const float* h_dets,
                             const int32_t* h_counts, int32_t* h_out_count, int32_t* h_out_ids, float* h_out_joints,
                             uint8_t* h_out_nviews, int32_t* h_out_assoc) {
    if (!h || !h_dets || !h_counts || !h_out_count || S <= 0 || T <= 0) return fail(h, PAM_E_INVALID, "bad argument");
    if (!h->have_cameras) return fail(h, PAM_E_NOCAMERAS, "pam_set_cameras has not been called");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    CK(cudaSetDevice(h->device));
    if (!h->ws_stream) CK(cudaStreamCreateWithFlags(&h->ws_stream, cudaStreamNonBlocking));
    cudaStream_t st = h->ws_stream;
    const DevCfg& c = h->dc;
    const size_t ST = (size_t)S * T;
    const size_t b_dets = ST * c.V * c.D * c.J * 3 * 4, b_counts = ST * c.V * 4, b_count = ST * 4;
    const size_t b_ids = ST * c.max_trk * 4, b_joints = ST * c.max_trk * c.J * 3 * 4, b_nv = ST * c.max_trk * c.J;
    const size_t b_assoc = ST * c.V * c.D * 4;
    if (fresh || S != h->ws_S) {
        CK(h->ws_state.reserve((size_t)c.seq_bytes * S));
        CK(cudaMemsetAsync(h->ws_state.p, 0, (size_t)c.seq_bytes * S, st));
        h->ws_S = S;
    }
    // Small jobs (the per-frame drop-in API: S = T = 1): no staging copies at all.  Inputs are placed in
    // pinned host memory that is mapped into the device address space; the kernel reads them and writes
    // its results over PCIe directly, so a call costs one launch and one stream synchronisation.
    {
        auto up = [](size_t x) { return (x + 255) / 256 * 256; };
        const size_t o_dets = 0, o_counts = o_dets + up(b_dets), o_count = o_counts + up(b_counts);
        const size_t o_ids = o_count + up(b_count), o_joints = o_ids + up(b_ids), o_nv = o_joints + up(b_joints);
        const size_t o_assoc = o_nv + up(b_nv), o_status = o_assoc + up(b_assoc), total = o_status + up((size_t)S * 4);
        if (total <= 256 * 1024) {
            if (total > h->zc_cap) {
                if (h->zc) cudaFreeHost(h->zc);
                h->zc = nullptr; h->zc_cap = 0;
                CK(cudaHostAlloc((void**)&h->zc, 256 * 1024, cudaHostAllocMapped));
                h->zc_cap = 256 * 1024;
            }
            char* z = h->zc;
            char* dz = nullptr;
            CK(cudaHostGetDevicePointer((void**)&dz, z, 0));
            memcpy(z + o_dets, h_dets, b_dets);
            memcpy(z + o_counts, h_counts, b_counts);
            TrackIO io{(const float*)(dz + o_dets), (const int32_t*)(dz + o_counts), (int32_t*)(dz + o_count),
                       h_out_ids ? (int32_t*)(dz + o_ids) : nullptr, h_out_joints ? (float*)(dz + o_joints) : nullptr,
                       h_out_nviews ? (uint8_t*)(dz + o_nv) : nullptr, h_out_assoc ? (int32_t*)(dz + o_assoc) : nullptr, T,
                       (int*)(dz + o_status)};
            // completion is detected by polling the status words the CTAs write last (a few us sooner
            // than a stream synchronisation wakes up); bounded, then the stream is synchronised anyway
            volatile int32_t* stv = (volatile int32_t*)(z + o_status);
            const int32_t pending = 0x7fffffff;
            for (int s = 0; s < S; ++s) stv[s] = pending;
            int rc = launch_track(h, h->ws_state.p, S, T, frame0, io, st);
            if (rc != PAM_OK) return rc;
            {
                const auto t0 = std::chrono::steady_clock::now();
                const auto budget = std::chrono::microseconds(200 + 40 * (int64_t)T);
                bool done = false;
                for (int spin = 0; !done; ++spin) {
                    done = true;
                    for (int s = 0; s < S && done; ++s) done = stv[s] != pending;
                    if (!done && (spin & 63) == 63 && std::chrono::steady_clock::now() - t0 > budget) break;
                }
                std::atomic_thread_fence(std::memory_order_acquire);
                if (!done) CK(cudaStreamSynchronize(st));
            }
            memcpy(h_out_count, z + o_count, b_count);
            if (h_out_ids) memcpy(h_out_ids, z + o_ids, b_ids);
            if (h_out_joints) memcpy(h_out_joints, z + o_joints, b_joints);
            if (h_out_nviews) memcpy(h_out_nviews, z + o_nv, b_nv);
            if (h_out_assoc) memcpy(h_out_assoc, z + o_assoc, b_assoc);
            const int32_t* stw = (const int32_t*)(z + o_status);
            for (int s = 0; s < S; ++s)
                if (stw[s] != 0) return pam_track_status(h, h->ws_state.p, S, nullptr, st);   // formats the message
            return PAM_OK;
        }
    }
    CK(h->ws_dets.reserve(b_dets));
    CK(h->ws_counts.reserve(b_counts));
    CK(h->ws_count.reserve(b_count));
    if (h_out_ids) CK(h->ws_ids.reserve(b_ids));
    if (h_out_joints) CK(h->ws_joints.reserve(b_joints));
    if (h_out_nviews) CK(h->ws_nv.reserve(b_nv));
    if (h_out_assoc) CK(h->ws_assoc.reserve(b_assoc));
    // Pipeline over chunks of frames: the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap
    // the kernel of chunk k (three streams; tracker state stays in HBM between the chunk launches).
    if (!h->ws_in) CK(cudaStreamCreateWithFlags(&h->ws_in, cudaStreamNonBlocking));
    if (!h->ws_out) CK(cudaStreamCreateWithFlags(&h->ws_out, cudaStreamNonBlocking));
    // an error return below must not leave copies into / out of the caller's buffers in flight
    struct DrainOnError {
        pam_handle* h;
        bool armed;
        ~DrainOnError() {
            if (!armed) return;
            cudaStreamSynchronize(h->ws_in); cudaStreamSynchronize(h->ws_stream); cudaStreamSynchronize(h->ws_out);
        }
    } drain{h, true};
    int nchunks = T / 50;   // PCIe-bound: more, smaller chunks shorten the pipeline fill and drain
    const char* nce = getenv("PAM_HOST_CHUNKS");
    if (nce) nchunks = atoi(nce);
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 64) nchunks = 64;
    if (nchunks > T) nchunks = T;
    while ((int)h->ev_in.size() < nchunks + 1) {
        cudaEvent_t e1, e2;
        CK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        h->ev_in.push_back(e1); h->ev_k.push_back(e2);
    }
    // the first H2D copy must not overtake earlier work queued on the compute stream (state reset)
    CK(cudaEventRecord(h->ev_k[nchunks], st));
    CK(cudaStreamWaitEvent(h->ws_in, h->ev_k[nchunks], 0));
    const size_t f_dets = (size_t)c.V * c.D * c.J * 3 * 4, f_counts = (size_t)c.V * 4, f_count = 4;
    const size_t f_ids = (size_t)c.max_trk * 4, f_joints = (size_t)c.max_trk * c.J * 3 * 4, f_nv = (size_t)c.max_trk * c.J;
    const size_t f_assoc = (size_t)c.V * c.D * 4;
    auto chunk2d = [&](void* dst, const void* src, size_t fbytes, int t0, int n, cudaMemcpyKind kind, cudaStream_t sx) {
        return cudaMemcpy2DAsync((char*)dst + (size_t)t0 * fbytes, (size_t)T * fbytes, (const char*)src + (size_t)t0 * fbytes,
                                 (size_t)T * fbytes, (size_t)n * fbytes, (size_t)S, kind, sx);
    };
    for (int k = 0; k < nchunks; ++k) {
        const int t0 = (int)((int64_t)T * k / nchunks), t1 = (int)((int64_t)T * (k + 1) / nchunks), n = t1 - t0;
        CK(chunk2d(h->ws_dets.p, h_dets, f_dets, t0, n, cudaMemcpyHostToDevice, h->ws_in));
        CK(chunk2d(h->ws_counts.p, h_counts, f_counts, t0, n, cudaMemcpyHostToDevice, h->ws_in));
        CK(cudaEventRecord(h->ev_in[k], h->ws_in));
        CK(cudaStreamWaitEvent(st, h->ev_in[k], 0));
        TrackIO io{(const float*)h->ws_dets.p + (size_t)t0 * (f_dets / 4), (const int32_t*)h->ws_counts.p + (size_t)t0 * c.V,
                   (int32_t*)h->ws_count.p + t0,
                   h_out_ids ? (int32_t*)h->ws_ids.p + (size_t)t0 * c.max_trk : nullptr,
                   h_out_joints ? (float*)h->ws_joints.p + (size_t)t0 * (f_joints / 4) : nullptr,
                   h_out_nviews ? (uint8_t*)h->ws_nv.p + (size_t)t0 * f_nv : nullptr,
                   h_out_assoc ? (int32_t*)h->ws_assoc.p + (size_t)t0 * (f_assoc / 4) : nullptr, T, nullptr};
        int rc = launch_track(h, h->ws_state.p, S, n, frame0 + t0, io, st);
        if (rc != PAM_OK) return rc;
        CK(cudaEventRecord(h->ev_k[k], st));
        CK(cudaStreamWaitEvent(h->ws_out, h->ev_k[k], 0));
        CK(chunk2d(h_out_count, h->ws_count.p, f_count, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_ids) CK(chunk2d(h_out_ids, h->ws_ids.p, f_ids, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_joints) CK(chunk2d(h_out_joints, h->ws_joints.p, f_joints, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_nviews) CK(chunk2d(h_out_nviews, h->ws_nv.p, f_nv, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_assoc) CK(chunk2d(h_out_assoc, h->ws_assoc.p, f_assoc, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
    }
    CK(cudaStreamSynchronize(h->ws_out));
    drain.armed = false;
    return pam_track_status(h, h->ws_state.p, S, nullptr, st);
}

int pam_track_state_to_host(pam_handle* h, int32_t S, void* h_state) {
    if (!h || !h_state || S <= 0 || S > h->ws_S) return fail(h, PAM_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h_state, h->ws_state.p, (size_t)h->dc.seq_bytes * S, cudaMemcpyDeviceToHost, h->ws_stream));
    CK(cudaStreamSynchronize(h->ws_stream));
    return PAM_OK;
}

int64_t pam_launch_count(const pam_handle* h) { return h ? h->launches : 0; }

}  // extern "C"

#include "pam_ops_capi.inc"
