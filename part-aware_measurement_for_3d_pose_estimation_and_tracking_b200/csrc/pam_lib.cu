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
// persistent tracker kernel: one thread GROUP (G warps) owns one sequence for all T frames; Q groups
// (sequences) share a CTA and the camera constants it keeps in shared memory
// ----------------------------------------------------------------------------------------------
template <int G, int AFF_UNROLL>
struct GroupCtx {
    static constexpr int kAffinityUnroll = AFF_UNROLL;
    static constexpr bool kTwoPassAffinity = (G == 1);      // throughput launches only: it adds a barrier to the frame
    int t;        // thread index inside the group
    int bar;      // named barrier of the group (1..15), unused for one-warp groups
    __device__ __forceinline__ int tid() const { return t; }
    __device__ __forceinline__ int nthreads() const { return G * 32; }
    __device__ __forceinline__ void sync() const {
        if (G == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(G * 32) : "memory");
    }
    // the seven barriers every frame passes; convoy level 2 makes them span the sequences of the CTA, so that all
    // its warps walk the frame's code together (instruction-cache locality at the price of waiting for the slowest)
    int convoy_threads;
    __device__ __forceinline__ void phase_sync() const {
        if (convoy_threads) asm volatile("bar.sync 0, %0;" ::"r"(convoy_threads) : "memory");
        else sync();
    }
    __device__ __forceinline__ void atomic_inc(int* p) const { atomicAdd(p, 1); }
    __device__ __forceinline__ int atomic_inc_ret(int* p) const { return atomicAdd(p, 1); }
    __device__ __forceinline__ void atomic_max_u64(unsigned long long* p, unsigned long long v) const { atomicMax(p, v); }
    __device__ __forceinline__ void atomic_min(int* p, int v) const { atomicMin(p, v); }
    __device__ __forceinline__ void atomic_min_u64(unsigned long long* p, unsigned long long v) const { atomicMin(p, v); }
    __device__ __forceinline__ int enum_limit() const { return 1024 * G; }     // candidate assignments worth enumerating: <= 32 per thread
    __device__ __forceinline__ long long clock() const { return clock64(); }
};

// Asynchronous global -> shared staging of one frame's detections.  When the frame is a whole number of
// 16-byte units (Shelf: 3360 B) ONE thread issues ONE bulk copy through the TMA engine
// (cp.async.bulk, completion counted in bytes on an mbarrier -- UBLKCP in SASS); otherwise all threads of
// the group issue 4-byte LDGSTS copies.  The V per-camera counts always go through LDGSTS.
// Two buffers: frame t+1 is staged while frame t is processed.  One buffer (throughput launches, half the
// shared memory): the copy of frame t+1 starts as soon as frame t no longer needs its detections (after the
// part-aware filter), and the many other sequences of the SM cover its latency.
__device__ __forceinline__ bool stage_is_bulk(const float* gsrc, int nfloats) {
    return (nfloats & 3) == 0 && ((uintptr_t)gsrc & 15) == 0;
}
__device__ __forceinline__ void stage_frame(int t, int nt, float* sdst, int* scnt, const float* gsrc, const int* gcnt,
                                            int nfloats, int V, unsigned long long* mbar, bool bulk) {
    if (bulk) {
        if (t == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(sdst), bar = (unsigned)__cvta_generic_to_shared(mbar);
            const unsigned bytes = (unsigned)nfloats * 4u;
            // order the generic-proxy reads of this buffer before the async-proxy write
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
        }
    } else {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sdst);
        PAM_NOUNROLL for (int i = t; i < nfloats; i += nt)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sbase + i * 4), "l"(gsrc + i) : "memory");
    }
    const unsigned cbase = (unsigned)__cvta_generic_to_shared(scnt);
    PAM_NOUNROLL for (int i = t; i < V; i += nt)
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

struct TrackIO {
    const float* dets;      // [S][T][V][D][J][3]
    const int* counts;      // [S][T][V]
    int* out_count;         // [S][T]
    int* out_ids;           // [S][T][MT]
    float* out_joints;      // [S][T][MT][J][3]
    unsigned char* out_nv;  // [S][T][MT][J]
    int* out_assoc;         // [S][T][V][D]
    int* out_timing;        // [S][T][4]
    unsigned char* out_vlist;  // [S][T][MT][PAM_VLIST]
    int seq_frames;         // frames between consecutive sequences in every tensor above (>= T)
    int* out_status;        // [S] final status word of each sequence, or null
};

// single-buffer staging: frame_step tells us when the staged detections are dead.  The hook only carries the
// index of the next frame; everything else is recomputed from values the kernel keeps anyway.
struct StageHook {
    int next;               // frame (inside the launch) to stage, -1 = nothing to do
    int t, nt;
    float* dbuf; int* cbuf; const float* gd; const int* gc; int nfl, V; unsigned long long* mbar; bool bulk;
    __device__ __forceinline__ void dets_released() const {
        if (next >= 0) stage_frame(t, nt, dbuf, cbuf, gd + (int64_t)next * nfl, gc + next * V, nfl, V, mbar, bulk);
    }
};

// dynamic shared memory of k_track_sequences: [camera constants][Q x sequence arena]
static inline size_t track_smem_bytes(const DevCfg& c, int Q) {
    return (size_t)cam_bytes_of(c) + (size_t)Q * c.arena_bytes;
}

#define PAM_TRACK_THREADS_MAX 1024

template <class K, int MAXT, int MINB, int G>
__global__ void __launch_bounds__(MAXT, MINB)
k_track_sequences(const DevCfg c, const CamConst cc, char* __restrict__ state, int S, int T, int frame0, const TrackIO io) {
    extern __shared__ __align__(128) char smem[];
    CamShared<K>* cam = (CamShared<K>*)smem;
    constexpr int NT = G * 32;
    constexpr int CAMB = (int)((sizeof(CamShared<K>) + 127) / 128 * 128);
    // warp-uniform by construction (a group is made of whole warps); saying so lets the compiler keep the
    // group's base addresses in uniform registers instead of recomputing them from threadIdx all over the frame
    const int grp = (MAXT == NT) ? 0 : __shfl_sync(0xffffffffu, (int)threadIdx.x / NT, 0);
    const int Q = (MAXT == NT) ? 1 : (int)blockDim.x / NT;
    GroupCtx<G, (MAXT * MINB <= 512) ? 4 : 2> ctx{(int)threadIdx.x - grp * NT, 1 + grp, 0};   // <= 512 resident threads per SM: 128 registers each
    char* arena = smem + CAMB + grp * c.arena_bytes;
    load_cameras(threadIdx.x, blockDim.x, c, cam, cc);
    SeqShared<K>& sh = *(SeqShared<K>*)arena;
    if (ctx.tid() == 0) {
        const unsigned b0 = (unsigned)__cvta_generic_to_shared(&sh.mbar[0]), b1 = (unsigned)__cvta_generic_to_shared(&sh.mbar[1]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int s = blockIdx.x * Q + grp;
    if (s >= S) return;                 // whole groups leave; every later barrier is per group
    Seq<K> sq;
    sq.bind(c, arena, cam, state + (int64_t)s * c.seq_bytes);
    const int nfl = c.V * c.D * c.J * 3;
    const int nfl_pad = c.frame_floats;
    const int nbuf = c.nbuf;
    float* dbuf = (float*)(arena + c.a_dets);
    int* cbuf = &sh.cnt[0][0];
    const float* gd = io.dets + (int64_t)s * io.seq_frames * nfl;
    const int* gc = io.counts + (int64_t)s * io.seq_frames * c.V;
    const bool bulk = stage_is_bulk(gd, nfl);
    if (T > 0) stage_frame(ctx.tid(), NT, dbuf, cbuf, gd, gc, nfl, c.V, &sh.mbar[0], bulk);
    load_state(ctx, c, sq);
    if (T > 0) stage_wait(&sh.mbar[0], 0u, bulk);
    ctx.sync();
    FrameOut o;
    {
        const int64_t f0 = (int64_t)s * io.seq_frames;
        o.count = io.out_count ? io.out_count + f0 : nullptr;
        o.ids = io.out_ids ? io.out_ids + f0 * c.max_rep : nullptr;
        o.joints = io.out_joints ? io.out_joints + f0 * c.max_rep * c.J * 3 : nullptr;
        o.nviews = io.out_nv ? io.out_nv + f0 * c.max_rep * c.J : nullptr;
        o.assoc = io.out_assoc ? io.out_assoc + f0 * c.V * c.D : nullptr;
        o.timing = io.out_timing ? io.out_timing + f0 * 4 : nullptr;
        o.vlist = io.out_vlist ? io.out_vlist + f0 * c.max_rep * PAM_VLIST : nullptr;
    }
    const int st_ids = c.max_rep, st_joints = c.max_rep * c.J * 3, st_nv = c.max_rep * c.J, st_assoc = c.V * c.D;
    StageHook hook{-1, ctx.tid(), NT, dbuf, cbuf, gd, gc, nfl, c.V, &sh.mbar[0], bulk};
    // convoy: the sequences of a CTA start every frame together, so their warps walk the same code at about the
    // same time and share instruction fetches (24 independent warps otherwise thrash the instruction caches)
    const int convoy_threads = (c.convoy && G == 1 && Q > 1) ? (((S - (int)blockIdx.x * Q) < Q ? (S - (int)blockIdx.x * Q) : Q) * NT) : 0;
    if (c.convoy >= 2 && G == 1 && convoy_threads > NT) ctx.convoy_threads = convoy_threads;
    PAM_NOUNROLL for (int t = 0; t < T; ++t) {
        if (convoy_threads > NT) asm volatile("bar.sync 0, %0;" ::"r"(convoy_threads) : "memory");
        const int cur = (nbuf == 2) ? (t & 1) : 0, nxt = (nbuf == 2) ? (cur ^ 1) : 0;
        const bool more = t + 1 < T;
        hook.next = (more && nbuf == 1) ? t + 1 : -1;
        if (more && nbuf == 2)
            stage_frame(ctx.tid(), NT, dbuf + nxt * nfl_pad, cbuf + nxt * PAM_MAX_V, gd + (int64_t)(t + 1) * nfl,
                        gc + (t + 1) * c.V, nfl, c.V, &sh.mbar[nxt], bulk);
        frame_step(ctx, c, sq, frame0 + t, dbuf + cur * nfl_pad, cbuf + cur * PAM_MAX_V, o, gd, frame0, hook);
        if (o.count) o.count += 1;
        if (o.ids) o.ids += st_ids;
        if (o.joints) o.joints += st_joints;
        if (o.nviews) o.nviews += st_nv;
        if (o.assoc) o.assoc += st_assoc;
        if (o.timing) o.timing += 4;
        if (o.vlist) o.vlist += st_ids * PAM_VLIST;
        // two buffers: buffer (t+1)&1 is used for the ((t+1)>>1)-th time; one buffer: for the (t+1)-th time --
        // that is the parity of its barrier phase
        if (more) stage_wait(&sh.mbar[nxt], (unsigned)((nbuf == 2 ? ((t + 1) >> 1) : (t + 1)) & 1), bulk);
        ctx.sync();
    }
    // the launch's last frame is still staged on chip when its buffer was not reused
    persist_views(ctx, c, sq, gd, frame0, T > 0 ? dbuf + ((nbuf == 2) ? ((T - 1) & 1) : 0) * nfl_pad : nullptr, T - 1);
    store_state(ctx, c, sq);
    if (io.out_status) {
        // the status word may live in mapped host memory and be polled by the caller (small-job path):
        // everything this group wrote for the host must be visible system-wide first, and nobody may still be
        // reading the caller's input buffer (persist_views) when the host sees completion
        ctx.sync();
        if (ctx.tid() == 0) {
            __threadfence_system();
            *(volatile int*)(io.out_status + s) = sh.hdr.status | (sh.hdr.warn << 8);
        }
    }
}

// ----------------------------------------------------------------------------------------------
// resident per-frame kernel ("stream mode"): the literal call pattern of the reference is ONE tracking() call
// per frame (src/ivclabpose.py:257).  A launch per call costs launch + state load/store + completion latency, so
// for that pattern one CTA stays resident: it polls a command word in pinned, device-mapped host memory, reads the
// frame's detections from the same slot over PCIe, runs frame_step with the tracker state kept in shared memory, and
// writes the results and a completion word back.  The host side of a call is: fill the slot, bump the command
// word, spin on the completion word.
// ----------------------------------------------------------------------------------------------
enum { STREAM_CMD_IDLE = 0, STREAM_CMD_EXIT = -1 };
struct StreamSlot {                      // offsets (bytes) into the mapped slot, computed by the host
    int o_cmd, o_frame, o_counts, o_dets, o_done, o_count, o_ids, o_joints, o_nv, o_assoc, o_timing, o_status, o_vlist, bytes;
};

template <class K, int G>
__global__ void __launch_bounds__(G * 32, 1)
k_track_stream(const DevCfg c, const CamConst cc, char* __restrict__ state, char* __restrict__ slot, const StreamSlot so,
               long long idle_limit_cycles) {
    extern __shared__ __align__(128) char smem[];
    CamShared<K>* cam = (CamShared<K>*)smem;
    constexpr int NT = G * 32;
    constexpr int CAMB = (int)((sizeof(CamShared<K>) + 127) / 128 * 128);
    GroupCtx<G, 4> ctx{(int)threadIdx.x, 1, 0};
    char* arena = smem + CAMB;
    load_cameras(threadIdx.x, NT, c, cam, cc);
    SeqShared<K>& sh = *(SeqShared<K>*)arena;
    __shared__ int s_cmd, s_frame, s_tail;
    __shared__ unsigned s_cnt8[2];
    __syncthreads();
    Seq<K> sq;
    sq.bind(c, arena, cam, state);
    load_state(ctx, c, sq);
    float* dbuf = (float*)(arena + c.a_dets);
    int* cbuf = &sh.cnt[0][0];
    const char* h_cmd = slot + so.o_cmd;   // 16 bytes: sequence number, frame id, per-camera counts (one byte each)
    FrameOut o;
    o.count = (int*)(slot + so.o_count); o.ids = (int*)(slot + so.o_ids); o.joints = (float*)(slot + so.o_joints);
    o.nviews = (unsigned char*)(slot + so.o_nv); o.assoc = (int*)(slot + so.o_assoc); o.timing = (int*)(slot + so.o_timing);
    o.vlist = (unsigned char*)(slot + so.o_vlist);
    NoHook hook;
    int last_seq = 0;
    ctx.sync();
    for (;;) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            int cmd;
            unsigned w0, w1, w2, w3;
            for (;;) {
                // one 16-byte PCIe read: sequence number, frame id and the per-camera counts together
                asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "l"(h_cmd) : "memory");
                cmd = (int)w0;
                if (cmd != last_seq) break;
                if (clock64() - t0 > idle_limit_cycles) { cmd = STREAM_CMD_EXIT; break; }    // nobody is calling: free the SM
            }
            s_cmd = cmd;
            s_frame = (int)w1;
            s_cnt8[0] = w2; s_cnt8[1] = w3;
        }
        ctx.sync();
        const long long tq0 = clock64();
        const int cmd = s_cmd;
        if (cmd == STREAM_CMD_EXIT) break;
        last_seq = cmd;
        const int frame = s_frame;
        // this frame's detections: host memory -> shared memory.  The host packs the detections of all cameras back
        // to back ([sum of counts][J][3], what one numpy.concatenate produces); the rows are spread to the padded
        // [V][D][J][3] layout here.  Uncached loads (ld.global.cv: the slot is rewritten between frames), eight in
        // flight per thread so that the whole frame costs ONE PCIe round trip.
        {
            int off[PAM_MAX_V + 1];
            int acc = 0;
#pragma unroll
            for (int v = 0; v < PAM_MAX_V; ++v) {
                off[v] = acc;
                int m = (v < c.V) ? (int)((s_cnt8[v >> 2] >> ((v & 3) * 8)) & 0xffu) : 0;
                m = m > c.D ? c.D : m;
                if (threadIdx.x == v) cbuf[v] = m;
                acc += m;
            }
            off[PAM_MAX_V] = acc;
            const int row = c.J * 3;                          // floats per detection
            const int total = acc * row;
            const float inv_row = 1.0f / (float)row;
            const float* src = (const float*)(slot + so.o_dets);
            PAM_NOUNROLL for (int base = 0; base < total; base += NT * 8) {
                float v8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + k * NT + (int)threadIdx.x;
                    v8[k] = (i < total) ? __ldcv(src + i) : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = base + k * NT + (int)threadIdx.x;
                    if (i >= total) continue;
                    const int r = fast_div(i, inv_row), e = i - r * row;
                    int v = 0;
#pragma unroll
                    for (int q = 1; q < PAM_MAX_V; ++q) v += (r >= off[q] && q < c.V) ? 1 : 0;
                    dbuf[((v * c.D) + (r - off[v])) * row + e] = v8[k];
                }
            }
        }
        ctx.sync();
        // stale views are read from the persisted copies (gin_frame0 = frame: only this frame's views count as
        // "inside the launch"), which persist_views refreshes after every frame
        const long long tq1 = clock64();
        frame_step(ctx, c, sq, frame, dbuf, cbuf, o, dbuf, frame, hook);
        ctx.sync();
        const long long tq2 = clock64();
        if (threadIdx.x == 0) {
            *(volatile int*)(slot + so.o_status) = sh.hdr.status | (sh.hdr.warn << 8);
            // protocol timing (SM cycles): input transfer, frame, and -- one frame late -- result write-back + persist
            int* tm = (int*)(slot + so.o_timing);
            tm[4] = (int)(tq1 - tq0); tm[5] = (int)(tq2 - tq1); tm[6] = s_tail;
        }
        // results first (every writer fences its own stores), then the completion word; the views of this frame
        // are persisted while the host already reads the results
        __threadfence_system();
        ctx.sync();
        if (threadIdx.x == 0) *(volatile int*)(slot + so.o_done) = cmd;
        const long long tq3 = clock64();
        persist_views(ctx, c, sq, dbuf, frame, dbuf, 0);
        ctx.sync();
        if (threadIdx.x == 0) s_tail = (int)(tq3 - tq2) | ((int)(clock64() - tq3) << 16);
    }
    store_state(ctx, c, sq);
    ctx.sync();
    if (threadIdx.x == 0) {
        __threadfence_system();
        *(volatile int*)(slot + so.o_done) = STREAM_CMD_EXIT;
    }
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
    DevCfg dc;               // latency flavour of the working set: two detection buffers, raw pose in the arena
    DevCfg dc_tp;            // throughput flavour: one detection buffer, raw pose in the sequence's HBM scratch
    int device = 0;
    bool have_cameras = false;
    int num_sms = 148;
    int smem_optin = 227 * 1024;
    int clock_khz = 0;
    // PAM_TRACK_SHAPE="<G>:<Q>:<regs>[:lean]" forces the launch shape (development / tests)
    int force_g = 0, force_q = 0, force_regs = 0, force_lean = -1;
    DevBuf cam;              // packed camera constants
    CamConst cc{};
    // workspace of the *_host entry points
    DevBuf ws_state, ws_dets, ws_counts, ws_count, ws_ids, ws_joints, ws_nv, ws_assoc, ws_timing, ws_vlist;
    int ws_S = 0;
    cudaStream_t ws_stream = nullptr, ws_in = nullptr, ws_out = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    // stream mode (resident per-frame kernel)
    char* st_slot = nullptr;   // pinned, device-mapped command / result slot
    char* st_slot_dev = nullptr;
    StreamSlot st_so{};
    bool st_running = false;
    int st_seq = 0, st_frame = 0;
    cudaStream_t st_stream = nullptr;
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
        case PAM_E_CAPACITY: return "a sequence exceeded a capacity limit (tracks / hypotheses / detections): the excess was dropped";
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
    if (tracker_capable(*cfg)) {
        rc = make_devcfg(*cfg, h->dc_tp, err, 1, false);
        if (rc != PAM_OK) { delete h; return fail(nullptr, rc, err); }
    } else {
        h->dc_tp = h->dc;
    }
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device);
    {
        // default 1: the one-warp sequences of a CTA start every frame together (+10 % on B200: the 16-24
        // independent warps of an SM otherwise thrash the instruction caches); 2: every phase together; 0: free running
        const char* cv = getenv("PAM_TRACK_CONVOY");
        const int convoy = cv ? atoi(cv) : 1;
        h->dc.convoy = convoy; h->dc_tp.convoy = convoy;
    }
    const char* shp = getenv("PAM_TRACK_SHAPE");
    if (shp) {
        int g = 0, q = 0, r = 0, lean = -1;
        const int got = sscanf(shp, "%d:%d:%d:%d", &g, &q, &r, &lean);
        if (got >= 1) h->force_g = g;
        if (got >= 2) h->force_q = q;
        if (got >= 3) h->force_regs = r;
        if (got >= 4) h->force_lean = lean;
    }
    *out = h;
    return PAM_OK;
}

int pam_destroy(pam_handle* h) {
    if (!h) return PAM_OK;
    cudaSetDevice(h->device);
    pam_stream_close(h);
    if (h->st_slot) cudaFreeHost(h->st_slot);
    if (h->st_stream) cudaStreamDestroy(h->st_stream);
    h->cam.release();
    h->ws_state.release(); h->ws_dets.release(); h->ws_counts.release(); h->ws_count.release();
    h->ws_ids.release(); h->ws_joints.release(); h->ws_nv.release(); h->ws_assoc.release(); h->ws_timing.release(); h->ws_vlist.release();
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

typedef void (*track_kernel_t)(const DevCfg, const CamConst, char*, int, int, int, const TrackIO);

// Launch shapes.  The kernel source is one; the variants differ in the register budget (__launch_bounds__) and
// in the warps per sequence G.  Q (sequences per CTA) is a run-time value: blockDim = Q * G * 32 <= maxt.
struct TrackVariant {
    int caps, maxt, minb, g, regs;
    track_kernel_t fn;
};
#define PAM_VARIANT(K, caps, maxt, minb, g, regs) {caps, maxt, minb, g, regs, k_track_sequences<K, maxt, minb, g>}
static const TrackVariant k_variants[] = {
    // ---- Campus / Shelf shaped working sets -------------------------------------------------------
    // one warp per sequence: throughput launches, 16-24 sequences per SM (measured on B200, Shelf shape,
    // frames of a CTA started together: 67.7 M frames/s at 16 per SM with 128 registers, 68.9 M at 24 per SM
    // with 80; without the convoy barrier 61.5 M; three warps per sequence at 8 per SM: 59 M)
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 256, 2, 1, 128),
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 320, 2, 1, 96),
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 384, 2, 1, 80),
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 256, 3, 1, 80),
    // three warps per sequence, one sequence per CTA, 8 CTAs per SM: 5-11 sequences per SM
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 96, 8, 3, 80),
    // four / eight warps per sequence: few sequences, latency matters, full register budget
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 128, 4, 4, 128),
    PAM_VARIANT(CapsSmall, CAPS_SMALL, 256, 2, 8, 128),
#if !defined(PAM_DEV_SMALL_ONLY)
    // ---- Panoptic shaped ---------------------------------------------------------------------------
    PAM_VARIANT(CapsMid, CAPS_MID, 192, 2, 1, 128),
    PAM_VARIANT(CapsMid, CAPS_MID, 256, 3, 1, 80),
    PAM_VARIANT(CapsMid, CAPS_MID, 96, 8, 3, 80),
    PAM_VARIANT(CapsMid, CAPS_MID, 128, 4, 4, 128),
    PAM_VARIANT(CapsMid, CAPS_MID, 256, 2, 8, 128),
    // ---- anything the tracker accepts (8 cameras x 16 detections x 32 joints x 32 tracks) ------------
    PAM_VARIANT(CapsMax, CAPS_MAX, 256, 2, 1, 128),
    PAM_VARIANT(CapsMax, CAPS_MAX, 128, 4, 4, 128),
    PAM_VARIANT(CapsMax, CAPS_MAX, 256, 2, 8, 128),
#endif
};
static const int k_num_variants = (int)(sizeof(k_variants) / sizeof(k_variants[0]));

struct TrackLaunch {
    const TrackVariant* v = nullptr;
    const DevCfg* cfg = nullptr;
    int q = 1, threads = 0, ctas_per_sm = 0;
    size_t smem = 0;
};

// CTAs of this shape one SM can hold (shared memory, threads, register-bound minb)
static int resident_ctas(const pam_handle* h, const TrackVariant& v, const DevCfg& c, int q) {
    const size_t smem = track_smem_bytes(c, q) + 1024;           // + the per-CTA reservation of the driver
    if (track_smem_bytes(c, q) > (size_t)h->smem_optin) return 0;
    int by_smem = (int)((size_t)228 * 1024 / smem);
    int by_threads = 2048 / (q * v.g * 32);
    int n = by_smem < by_threads ? by_smem : by_threads;
    if (n > v.minb) n = v.minb;
    if (n > 32) n = 32;
    return n;
}

// Launch shape per call.  Few sequences per SM: latency matters -- many warps per sequence, full register
// budget, two detection buffers.  Many sequences per SM: one warp per sequence, as many sequences resident as
// the shared memory allows (no block barriers, nobody waits for a neighbour's phase), one detection buffer.
static TrackLaunch pick_track_launch_g(const pam_handle* h, int S, int g, bool lean) {
    TrackLaunch L;
    const int per_sm = (S + h->num_sms - 1) / h->num_sms;
    L.cfg = lean ? &h->dc_tp : &h->dc;
    int best = -1, best_res = -1;
    for (int k = 0; k < k_num_variants; ++k) {
        const TrackVariant& v = k_variants[k];
        if (v.caps != L.cfg->caps || v.g != g) continue;
        if (h->force_regs && v.regs != h->force_regs) continue;
        int q = v.maxt / (g * 32);
        if (h->force_q) q = h->force_q;
        if (q * g * 32 > v.maxt || (g > 1 && q > 15)) continue;
        // not more groups per CTA than the batch can fill on every SM
        while (q > 1 && !h->force_q && (S + q - 1) / q < h->num_sms) --q;
        // ... nor than fit into shared memory
        while (q > 1 && !h->force_q && resident_ctas(h, v, *L.cfg, q) <= 0) --q;
        const int ctas = resident_ctas(h, v, *L.cfg, q);
        if (ctas <= 0) continue;
        int res = ctas * q;
        if (res >= per_sm && !h->force_regs) {
            // room for the whole batch: among those variants prefer the larger register budget
            res = per_sm * 1000 + v.regs;
        }
        if (res > best_res) { best_res = res; best = k; L.q = q; L.ctas_per_sm = ctas; }
    }
    if (best < 0) return L;
    L.v = &k_variants[best];
    L.threads = L.q * g * 32;
    L.smem = track_smem_bytes(*L.cfg, L.q);
    return L;
}

static TrackLaunch pick_track_launch(const pam_handle* h, int S) {
    const int per_sm = (S + h->num_sms - 1) / h->num_sms;
    if (h->force_g) return pick_track_launch_g(h, S, h->force_g, h->force_lean >= 0 ? h->force_lean != 0 : h->force_g == 1);
    // measured (B200, Shelf shape): <= 4 sequences per SM four warps each (44.7 M frames/s at 4 per SM against 42.0 M
    // with three), 5-11 three warps each, from 12 per SM one warp each
    const int want = per_sm <= 4 ? 4 : (per_sm <= 11 ? 3 : 1);
    // not every capacity class has every group size: take the nearest one that exists and fits
    static const int order[4][5] = {{1, 2, 3, 4, 8}, {2, 1, 3, 4, 8}, {3, 2, 4, 1, 8}, {4, 8, 3, 2, 1}};
    const int* ord = order[want == 1 ? 0 : (want == 2 ? 1 : (want == 3 ? 2 : 3))];
    for (int k = 0; k < 5; ++k) {
        const int g = ord[k];
        const bool lean = h->force_lean >= 0 ? (h->force_lean != 0) : (g == 1);
        TrackLaunch L = pick_track_launch_g(h, S, g, lean);
        if (L.v) return L;
    }
    return TrackLaunch();
}

static int launch_track(pam_handle* h, void* d_state, int32_t S, int32_t T, int32_t frame0, const TrackIO& io,
                        cudaStream_t stream) {
    const TrackLaunch L = pick_track_launch(h, S);
    if (!L.v) return fail(h, PAM_E_INVALID, "no launch shape fits this configuration into shared memory (PAM_TRACK_SHAPE?)");
    CK(cudaFuncSetAttribute(L.v->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    const int grid = (S + L.q - 1) / L.q;
    L.v->fn<<<grid, L.threads, L.smem, stream>>>(*L.cfg, h->cc, (char*)d_state, S, T, frame0, io);
    h->launches += 1;
    CK(cudaGetLastError());
    return PAM_OK;
}

int pam_track_launch_info(pam_handle* h, int32_t S, int32_t* out8) {
    if (!h || !out8 || S <= 0) return fail(h, PAM_E_INVALID, "bad argument");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    const TrackLaunch L = pick_track_launch(h, S);
    if (!L.v) return fail(h, PAM_E_INVALID, "no launch shape fits this configuration into shared memory");
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, L.v->fn));
    out8[0] = L.v->g; out8[1] = L.q; out8[2] = L.threads; out8[3] = L.ctas_per_sm; out8[4] = (int32_t)L.smem;
    out8[5] = fa.numRegs; out8[6] = L.cfg->arena_bytes; out8[7] = L.cfg->nbuf;
    return PAM_OK;
}

int pam_sm_clock_khz(pam_handle* h, int32_t* khz) {
    if (!h || !khz) return PAM_E_INVALID;
    *khz = h->clock_khz;
    return PAM_OK;
}

int pam_track_sequences(pam_handle* h, void* d_state, int32_t S, int32_t T, int32_t frame0, const float* d_dets,
                        const int32_t* d_counts, int32_t* d_out_count, int32_t* d_out_ids, float* d_out_joints,
                        uint8_t* d_out_nviews, int32_t* d_out_assoc, int32_t* d_out_timing, uint8_t* d_out_vlist,
                        void* stream) {
    if (!h || !d_state || !d_dets || !d_counts || S < 0 || T < 0) return fail(h, PAM_E_INVALID, "bad argument");
    if (!h->have_cameras) return fail(h, PAM_E_NOCAMERAS, "pam_set_cameras has not been called");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    if (S == 0 || T == 0) return PAM_OK;
    CK(cudaSetDevice(h->device));
    TrackIO io{d_dets, d_counts, d_out_count, d_out_ids, d_out_joints, d_out_nviews, d_out_assoc, d_out_timing, d_out_vlist,
               T, nullptr};
    return launch_track(h, d_state, S, T, frame0, io, (cudaStream_t)stream);
}

static int status_message(pam_handle* h, const int32_t* st, int S) {
    int hard = 0, warned = 0, first = -1, code = 0;
    for (int s = 0; s < S; ++s) {
        if (st[s] & 0xff) { if (!hard) { first = s; code = st[s] & 0xff; } ++hard; }
        else if (st[s]) { if (!hard && !warned) { first = s; code = st[s] >> 8; } ++warned; }
    }
    char msg[320];
    if (hard) {
        snprintf(msg, sizeof msg, "%d sequence(s) stopped with an internal error; first: sequence %d, code %d%s", hard, first,
                 code, code == SEQ_ERR_HIST_OVERFLOW ? " (pose history ring exhausted)" : "");
        return fail(h, PAM_E_INTERNAL, msg);
    }
    if (warned) {
        snprintf(msg, sizeof msg,
                 "%d sequence(s) hit a capacity limit (the excess was dropped for that frame, tracking went on); first: "
                 "sequence %d:%s%s%s", warned, first,
                 (code & WARN_TRACK_OVERFLOW) ? " track slots exhausted (raise max_tracks)" : "",
                 (code & WARN_HYP_OVERFLOW) ? " hypothesis table exhausted" : "",
                 (code & WARN_DET_OVERFLOW) ? " more detections than max_detections" : "");
        return fail(h, PAM_E_CAPACITY, msg);
    }
    return PAM_OK;
}

int pam_track_status(pam_handle* h, const void* d_state, int32_t S, int32_t* h_status, void* stream) {
    if (!h || !d_state || S < 0) return fail(h, PAM_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    std::vector<int32_t> tmp((size_t)S), wrn((size_t)S);
    const char* base = (const char*)d_state + h->dc.off_hdr;
    CK(cudaMemcpy2DAsync(tmp.data(), 4, base + offsetof(SeqHeader, status), (size_t)h->dc.seq_bytes, 4, (size_t)S,
                         cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaMemcpy2DAsync(wrn.data(), 4, base + offsetof(SeqHeader, warn), (size_t)h->dc.seq_bytes, 4, (size_t)S,
                         cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    for (int s = 0; s < S; ++s) {
        tmp[s] = (tmp[s] & 0xff) | (wrn[s] << 8);
        if (h_status) h_status[s] = tmp[s];
    }
    return status_message(h, tmp.data(), S);
}

int pam_track_host_status(pam_handle* h, int32_t S, int32_t* h_status) {
    if (!h || S <= 0 || S > h->ws_S || !h->ws_stream) return fail(h, PAM_E_INVALID, "bad argument");
    return pam_track_status(h, h->ws_state.p, S, h_status, h->ws_stream);
}

int pam_track_margins(pam_handle* h, const void* d_state, int32_t S, double* h_margins, void* stream) {
    if (!h || !d_state || !h_margins || S < 0) return fail(h, PAM_E_INVALID, "bad argument");
#if defined(PAM_MARGIN)
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy2DAsync(h_margins, 8 * MG_COUNT, (const char*)d_state + h->dc.off_margin, (size_t)h->dc.seq_bytes,
                         8 * MG_COUNT, (size_t)S, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return PAM_OK;
#else
    return fail(h, PAM_E_INVALID, "this libpam.so was built without -DPAM_MARGIN (tools/build_margin.sh)");
#endif
}

int pam_track_sequences_host(pam_handle* h, int32_t S, int32_t T, int32_t frame0, int32_t fresh, const float* h_dets,
                             const int32_t* h_counts, int32_t* h_out_count, int32_t* h_out_ids, float* h_out_joints,
                             uint8_t* h_out_nviews, int32_t* h_out_assoc, int32_t* h_out_timing, uint8_t* h_out_vlist) {
    if (!h || !h_dets || !h_counts || !h_out_count || S <= 0 || T <= 0) return fail(h, PAM_E_INVALID, "bad argument");
    if (!h->have_cameras) return fail(h, PAM_E_NOCAMERAS, "pam_set_cameras has not been called");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    CK(cudaSetDevice(h->device));
    if (h->st_running) { int rc = pam_stream_close(h); if (rc != PAM_OK) return rc; }
    if (!h->ws_stream) CK(cudaStreamCreateWithFlags(&h->ws_stream, cudaStreamNonBlocking));
    cudaStream_t st = h->ws_stream;
    const DevCfg& c = h->dc;
    const size_t ST = (size_t)S * T;
    const size_t b_dets = ST * c.V * c.D * c.J * 3 * 4, b_counts = ST * c.V * 4, b_count = ST * 4;
    const size_t b_ids = ST * c.max_rep * 4, b_joints = ST * c.max_rep * c.J * 3 * 4, b_nv = ST * c.max_rep * c.J;
    const size_t b_assoc = ST * c.V * c.D * 4, b_timing = ST * 16, b_vlist = ST * c.max_rep * PAM_VLIST;
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
        const size_t o_assoc = o_nv + up(b_nv), o_timing = o_assoc + up(b_assoc), o_vlist = o_timing + up(b_timing);
        const size_t o_status = o_vlist + up(b_vlist);
        const size_t total = o_status + up((size_t)S * 4);
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
                       h_out_nviews ? (uint8_t*)(dz + o_nv) : nullptr, h_out_assoc ? (int32_t*)(dz + o_assoc) : nullptr,
                       h_out_timing ? (int32_t*)(dz + o_timing) : nullptr, h_out_vlist ? (uint8_t*)(dz + o_vlist) : nullptr, T,
                       (int*)(dz + o_status)};
            // completion is detected by polling the status words the groups write last (a few us sooner
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
            if (h_out_timing) memcpy(h_out_timing, z + o_timing, b_timing);
            if (h_out_vlist) memcpy(h_out_vlist, z + o_vlist, b_vlist);
            const int32_t* stw = (const int32_t*)(z + o_status);
            for (int s = 0; s < S; ++s)
                if (stw[s] & 0xff) return status_message(h, stw, S);     // hard errors only; warnings: pam_track_host_status
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
    if (h_out_timing) CK(h->ws_timing.reserve(b_timing));
    if (h_out_vlist) CK(h->ws_vlist.reserve(b_vlist));
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
    const size_t f_ids = (size_t)c.max_rep * 4, f_joints = (size_t)c.max_rep * c.J * 3 * 4, f_nv = (size_t)c.max_rep * c.J;
    const size_t f_assoc = (size_t)c.V * c.D * 4, f_timing = 16, f_vlist = (size_t)c.max_rep * PAM_VLIST;
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
                   h_out_ids ? (int32_t*)h->ws_ids.p + (size_t)t0 * c.max_rep : nullptr,
                   h_out_joints ? (float*)h->ws_joints.p + (size_t)t0 * (f_joints / 4) : nullptr,
                   h_out_nviews ? (uint8_t*)h->ws_nv.p + (size_t)t0 * f_nv : nullptr,
                   h_out_assoc ? (int32_t*)h->ws_assoc.p + (size_t)t0 * (f_assoc / 4) : nullptr,
                   h_out_timing ? (int32_t*)h->ws_timing.p + (size_t)t0 * 4 : nullptr,
                   h_out_vlist ? (uint8_t*)h->ws_vlist.p + (size_t)t0 * f_vlist : nullptr, T, nullptr};
        int rc = launch_track(h, h->ws_state.p, S, n, frame0 + t0, io, st);
        if (rc != PAM_OK) return rc;
        CK(cudaEventRecord(h->ev_k[k], st));
        CK(cudaStreamWaitEvent(h->ws_out, h->ev_k[k], 0));
        CK(chunk2d(h_out_count, h->ws_count.p, f_count, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_ids) CK(chunk2d(h_out_ids, h->ws_ids.p, f_ids, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_joints) CK(chunk2d(h_out_joints, h->ws_joints.p, f_joints, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_nviews) CK(chunk2d(h_out_nviews, h->ws_nv.p, f_nv, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_assoc) CK(chunk2d(h_out_assoc, h->ws_assoc.p, f_assoc, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_timing) CK(chunk2d(h_out_timing, h->ws_timing.p, f_timing, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
        if (h_out_vlist) CK(chunk2d(h_out_vlist, h->ws_vlist.p, f_vlist, t0, n, cudaMemcpyDeviceToHost, h->ws_out));
    }
    CK(cudaStreamSynchronize(h->ws_out));
    drain.armed = false;
    {   // hard errors fail the call; capacity warnings are left to pam_track_host_status
        std::vector<int32_t> stv((size_t)S);
        int rc = pam_track_status(h, h->ws_state.p, S, stv.data(), st);
        if (rc == PAM_E_CAPACITY) rc = PAM_OK;
        return rc;
    }
}

int pam_track_state_to_host(pam_handle* h, int32_t S, void* h_state) {
    if (!h || !h_state || S <= 0 || S > h->ws_S) return fail(h, PAM_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    if (h->st_running) { int rc = pam_stream_close(h); if (rc != PAM_OK) return rc; }   // state lives on chip while it runs
    CK(cudaMemcpyAsync(h_state, h->ws_state.p, (size_t)h->dc.seq_bytes * S, cudaMemcpyDeviceToHost, h->ws_stream));
    CK(cudaStreamSynchronize(h->ws_stream));
    return PAM_OK;
}

// ---- stream mode ---------------------------------------------------------------------------------------
typedef void (*stream_kernel_t)(const DevCfg, const CamConst, char*, char*, const StreamSlot, long long);

static int stream_launch(pam_handle* h) {
    const DevCfg& c = h->dc;
    stream_kernel_t fn = c.caps == CAPS_SMALL ? k_track_stream<CapsSmall, 4>
                       : (c.caps == CAPS_MID ? k_track_stream<CapsMid, 4> : k_track_stream<CapsMax, 8>);
    const size_t smem = track_smem_bytes(c, 1);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *(volatile int*)(h->st_slot + h->st_so.o_done) = 0;
    *(volatile unsigned long long*)(h->st_slot + h->st_so.o_cmd) = 0ull;
    h->st_seq = 0;
    const long long idle = (long long)h->clock_khz * 1000;        // about one second of SM cycles
    fn<<<1, c.caps == CAPS_MAX ? 256 : 128, smem, h->st_stream>>>(c, h->cc, (char*)h->ws_state.p, h->st_slot_dev, h->st_so, idle);
    h->launches += 1;
    CK(cudaGetLastError());
    h->st_running = true;
    return PAM_OK;
}

int pam_stream_open(pam_handle* h, int32_t fresh) {
    if (!h) return fail(h, PAM_E_INVALID, "null handle");
    if (!h->have_cameras) return fail(h, PAM_E_NOCAMERAS, "pam_set_cameras has not been called");
    if (!tracker_capable(h->cfg)) return fail(h, PAM_E_INVALID, "the stateful tracker handles at most 8 cameras");
    CK(cudaSetDevice(h->device));
    if (h->st_running) { int rc = pam_stream_close(h); if (rc != PAM_OK) return rc; }
    if (!h->ws_stream) CK(cudaStreamCreateWithFlags(&h->ws_stream, cudaStreamNonBlocking));
    if (!h->st_stream) CK(cudaStreamCreateWithFlags(&h->st_stream, cudaStreamNonBlocking));
    const DevCfg& c = h->dc;
    if (fresh || h->ws_S != 1) {
        CK(h->ws_state.reserve((size_t)c.seq_bytes));
        CK(cudaMemsetAsync(h->ws_state.p, 0, (size_t)c.seq_bytes, h->ws_stream));
        CK(cudaStreamSynchronize(h->ws_stream));
        h->ws_S = 1;
    }
    if (!h->st_slot) {
        StreamSlot so;
        int o = 0;
        auto take = [&](size_t bytes) { const int at = o; o += (int)((bytes + 127) / 128 * 128); return at; };
        so.o_cmd = take(16); so.o_frame = so.o_cmd + 4; so.o_counts = take(4 * PAM_MAX_V);
        so.o_dets = take((size_t)c.V * c.D * c.J * 3 * 4);
        so.o_done = take(4); so.o_count = take(4); so.o_ids = take((size_t)c.max_rep * 4);
        so.o_joints = take((size_t)c.max_rep * c.J * 3 * 4); so.o_nv = take((size_t)c.max_rep * c.J);
        so.o_assoc = take((size_t)c.V * c.D * 4); so.o_timing = take(32); so.o_status = take(4);
        so.o_vlist = take((size_t)c.max_rep * PAM_VLIST);
        so.bytes = o;
        CK(cudaHostAlloc((void**)&h->st_slot, (size_t)o, cudaHostAllocMapped));
        memset(h->st_slot, 0, (size_t)o);
        CK(cudaHostGetDevicePointer((void**)&h->st_slot_dev, h->st_slot, 0));
        h->st_so = so;
    }
    return stream_launch(h);
}

int pam_stream_buffers(pam_handle* h, pam_stream_views* v) {
    if (!h || !v || !h->st_slot) return fail(h, PAM_E_INVALID, "stream not open");
    const StreamSlot& so = h->st_so;
    v->dets = (float*)(h->st_slot + so.o_dets); v->counts = (int32_t*)(h->st_slot + so.o_counts);
    v->out_count = (int32_t*)(h->st_slot + so.o_count); v->out_ids = (int32_t*)(h->st_slot + so.o_ids);
    v->out_joints = (float*)(h->st_slot + so.o_joints); v->out_nviews = (uint8_t*)(h->st_slot + so.o_nv);
    v->out_assoc = (int32_t*)(h->st_slot + so.o_assoc); v->out_timing = (int32_t*)(h->st_slot + so.o_timing);
    v->out_status = (int32_t*)(h->st_slot + so.o_status);
    v->out_vlist = (uint8_t*)(h->st_slot + so.o_vlist);
    return PAM_OK;
}

int pam_stream_submit(pam_handle* h, int32_t frame_id) {
    if (!h || !h->st_slot) return fail(h, PAM_E_INVALID, "stream not open");
    volatile int* done = (volatile int*)(h->st_slot + h->st_so.o_done);
    if (!h->st_running || *done == STREAM_CMD_EXIT) {        // the kernel left after an idle second: start it again
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->st_stream));
        h->st_running = false;
        int rc = stream_launch(h);
        if (rc != PAM_OK) return rc;
    }
    int seq = h->st_seq + 1;
    if (seq <= 0) seq = 1;
    h->st_seq = seq;
    h->st_frame = frame_id;
    // the command line (16 bytes, one PCIe read for the kernel): per-camera counts as bytes first, then sequence
    // number + frame id in ONE 8-byte store
    {
        const int32_t* cnt = (const int32_t*)(h->st_slot + h->st_so.o_counts);
        unsigned long long packed = 0ull;
        for (int v = 0; v < h->dc.V; ++v) {
            int m = cnt[v];
            m = m < 0 ? 0 : (m > 255 ? 255 : m);
            packed |= (unsigned long long)m << (8 * v);
        }
        *(volatile unsigned long long*)(h->st_slot + h->st_so.o_cmd + 8) = packed;
    }
    std::atomic_thread_fence(std::memory_order_release);     // inputs before the command word
    *(volatile unsigned long long*)(h->st_slot + h->st_so.o_cmd) =
        (unsigned long long)(unsigned)seq | ((unsigned long long)(unsigned)frame_id << 32);
    return PAM_OK;
}

int pam_stream_wait(pam_handle* h) {
    if (!h || !h->st_slot) return fail(h, PAM_E_INVALID, "stream not open");
    volatile int* done = (volatile int*)(h->st_slot + h->st_so.o_done);
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int seq = h->st_seq;
        const auto t0 = std::chrono::steady_clock::now();
        for (int spin = 0;; ++spin) {
            const int d = *done;
            if (d == seq) {
                std::atomic_thread_fence(std::memory_order_acquire);
                const int st = *(volatile int*)(h->st_slot + h->st_so.o_status);
                if (st & 0xff) return status_message(h, &st, 1);
                return PAM_OK;
            }
            if (d == STREAM_CMD_EXIT) break;                     // raced with the idle exit: relaunch and resubmit
            if ((spin & 1023) == 1023) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) {
                    cudaError_t e = cudaStreamQuery(h->st_stream);
                    if (e != cudaErrorNotReady) { h->st_running = false; return cuda_fail(h, e == cudaSuccess ? cudaErrorUnknown : e, "resident kernel stopped"); }
                    return fail(h, PAM_E_INTERNAL, "resident kernel did not answer within 5 s");
                }
            }
        }
        int rc = pam_stream_submit(h, h->st_frame);              // the inputs are still in the slot
        if (rc != PAM_OK) return rc;
    }
    return fail(h, PAM_E_INTERNAL, "resident kernel keeps exiting");
}

int pam_stream_step(pam_handle* h, int32_t frame_id) {
    int rc = pam_stream_submit(h, frame_id);
    return rc != PAM_OK ? rc : pam_stream_wait(h);
}

int pam_stream_close(pam_handle* h) {
    if (!h) return PAM_OK;
    if (!h->st_running) return PAM_OK;
    CK(cudaSetDevice(h->device));
    *(volatile unsigned long long*)(h->st_slot + h->st_so.o_cmd) = 0xffffffffull;      // STREAM_CMD_EXIT
    CK(cudaStreamSynchronize(h->st_stream));       // the kernel stores the tracker state before it leaves
    h->st_running = false;
    return PAM_OK;
}

int64_t pam_launch_count(const pam_handle* h) { return h ? h->launches : 0; }

}  // extern "C"

#include "pam_ops_capi.inc"
