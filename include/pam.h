/* pam.h -- C ABI of libpam.so, the B200 (sm_100a) implementation of the per-frame geometric hot
 * path of Part-Aware Measurement (cross-view association, part-aware view filtering, weighted DLT
 * triangulation, track life-cycle).
 *
 * The reference has no FFI: its boundary is a set of plain Python functions / classes
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces
 * (paths relative to /root/reference/src); INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns an int status: 0 = PAM_OK, negative = pam_status error
 *   - no exceptions / aborts cross the ABI; pam_last_error(h) gives a message
 *   - "d_" arguments are DEVICE pointers owned by the caller, "h_" arguments are HOST pointers;
 *     the library never frees caller memory
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); "d_" entry points are
 *     asynchronous on that stream, "h_"/"_host" entry points synchronise before returning
 *   - one handle per (device, host thread)
 *   - 2-D detections are (v, u, conf) = (row, col, confidence) triples like the reference's
 *     (ivclabpose.py:238-244); 3-D joints are (x, y, z) in world units
 */
#ifndef PAM_H
#define PAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAM_ABI_VERSION 2

/* capacity limits of the stateful tracker (pam_core.h).  Tracks / detections / joints / cameras are run-time
 * values of pam_config below these bounds: the working set of a sequence is sized from them. */
#define PAM_LIMIT_CAMERAS 8        /* stateful tracker; stateless ops: 32 */
#define PAM_LIMIT_TRACKS 32
#define PAM_LIMIT_DETECTIONS 16
#define PAM_LIMIT_JOINTS 32

typedef enum pam_status {
    PAM_OK = 0,
    PAM_E_INVALID = -1,        /* bad argument / configuration outside the limits above          */
    PAM_E_CUDA = -2,           /* a CUDA runtime call failed (no device, out of memory, ...)      */
    PAM_E_NOCAMERAS = -3,      /* pam_set_cameras has not been called                              */
    PAM_E_CAPACITY = -4,       /* a sequence exceeded max_tracks / hypotheses / max_detections: the excess
                                  was DROPPED for that frame and tracking went on (pam_track_status) */
    PAM_E_INTERNAL = -5
} pam_status;

/* Tracker hyper-parameters = the fields ivclabpose copies into iter_args (ivclabpose.py:140-156)
 * that live code reads, plus the constants the reference hard-wires. */
typedef struct pam_config {
    int32_t num_cameras;        /* V                                                              */
    int32_t num_joints;         /* J   (reference: 17 hard-wired)                                  */
    int32_t max_detections;     /* D   stride of the detection tensors                             */
    int32_t max_tracks;         /* track slots per sequence = stride of the output tensors        */
    int32_t n_init;             /* N_INIT   tracking/IterativeTracker.py:259                       */
    int32_t max_age;            /* MAX_AGE  :273, :331                                             */
    int32_t min_valid_joints;   /* the "> 10" of :145                                              */
    int32_t stale_window;       /* the "<= 3" of :317                                              */
    uint32_t arm_joint_mask;    /* bit j set: joint j is smoothed with arm_sigma (":382" [9,10])   */
    uint32_t max_report;        /* rows per frame of the output tensors (ids, joints, nviews, vlist); 0 = max_tracks.
                                   A frame that reports more tracks keeps the first max_report rows; d_out_count still
                                   holds the true number, so truncation is detectable.  Smaller rows = less D2H traffic */
    double conf_threshold;      /* CONF_THRESHOLD  :59                                             */
    double epi_threshold;       /* EPI_THRESHOLD   tracking/hypothesis.py:63                       */
    double init_threshold;      /* INIT_THRESHOLD  tracking/hypothesis.py:32                       */
    double joint_threshold;     /* JOINT_THRESHOLD :346                                            */
    double alpha2d;             /* ALPHA2D  :143                                                   */
    double lambda_a;            /* LAMBDA_A :148                                                   */
    double lambda_t;            /* LAMBDA_T utils/construction.py:96                               */
    double sigma;               /* SIGMA     :381                                                  */
    double arm_sigma;           /* ARM_SIGMA :382                                                  */
    double veto_believe;        /* the "> 0.5" of tracking/hypothesis.py:66                        */
} pam_config;

/* Byte layout of one sequence's tracker state inside the caller's state buffer, so that host
 * code can read tracks back (IterTrack read surface, tracking/IterativeTracker.py:194-204). */
typedef struct pam_state_layout {
    int64_t seq_bytes;          /* stride between sequences                                         */
    int64_t off_header;         /* int32 x header_ints: ntracks, next_id, status, frames_done, used_mask,
                                   warn (PAM_WARN_* bits), warn_frames, reserved; then int8 order[max_order] */
    int64_t off_meta;           /* per slot, int32 x meta_ints                                      */
    int64_t off_hist;           /* double [slot][hist_ring][J][3]  smoothed pose ring               */
    int64_t off_view;           /* float  [slot][V][J][3]         last (v,u,conf) per view slot;
                                   written at the END of every launch (inside a launch stale views
                                   are read from the launch's own detection tensor)               */
    int64_t off_vel;            /* float  [slot][J][3]            velocity                          */
    int64_t off_nviews;         /* uint8  [slot][J]               views used for the last pose      */
    int64_t off_margin;         /* double [n_margins]  decision margins (PAM_MARGIN builds only)    */
    int32_t meta_ints;          /* ints per slot: id,hits,age,tsu,state,already,nviews,hist_start,
                                   hist_len, vt_last (usable views of the last update), view_cid[8],
                                   view_time[8], hist_time[hist_ring],
                                   then 8 bytes camera -> view-slot map and 8 bytes
                                   detection index of each view inside its own frame             */
    int32_t hist_ring;          /* ring length                                                      */
    int32_t max_views;          /* 8                                                                */
    int32_t max_order;          /* 32                                                               */
    int32_t header_ints;        /* 8                                                                */
    int32_t n_margins;          /* 8                                                                */
} pam_state_layout;

typedef struct pam_handle pam_handle;

int pam_abi_version(void);
const char* pam_status_string(int status);
const char* pam_last_error(const pam_handle* h);

/* IterativeTracker.__init__ (tracking/IterativeTracker.py:36-45): validate + freeze parameters. */
int pam_create(const pam_config* cfg, int device, pam_handle** out);
int pam_destroy(pam_handle* h);

/* Camera constants of ivclabpose.GetCameraParameters / Camera.__init__ (ivclabpose.py:35-46,
 * 162-181), HOST pointers, copied to the device: P [V][3][4] f32, RKinv [V][3][3] f32,
 * position [V][3] f64, F [V][V][3][3] f32 with F[a][b] = cameras[a].F[b]. */
int pam_set_cameras(pam_handle* h, const float* h_P, const float* h_RKinv, const double* h_position,
                    const float* h_F);

/* ---- stateful path: IterativeTracker.tracking (tracking/IterativeTracker.py:115-180) ---------- */

int pam_get_state_layout(const pam_handle* h, pam_state_layout* out);

/* track_restart (tracking/IterativeTracker.py:47-50) for S sequences: d_state = S * seq_bytes. */
int pam_track_reset(pam_handle* h, void* d_state, int32_t S, void* stream);

/* Run T consecutive frames (frame ids frame0 .. frame0+T-1) of S independent sequences with no host
 * involvement between frames.  One thread group (1-8 warps, chosen from S) owns one sequence; several
 * groups share a CTA and its camera constants.
 *   d_dets   [S][T][V][D][J][3] f32 (v,u,conf), zero padded        d_counts [S][T][V] i32
 *   d_out_count [S][T] i32   number of reported tracks (Confirmed and updated this frame,
 *                            ivclabpose.py:265-267), in track-list order
 *   d_out_ids   [S][T][R] i32          d_out_joints [S][T][R][J][3] f32      (R = max_report, default max_tracks)
 *   d_out_nviews [S][T][R][J] u8 (may be NULL)  views each joint was built from
 *   d_out_assoc  [S][T][V][D] i32 (may be NULL) track id matched to each detection, -1 = none
 *   d_out_timing [S][T][4] i32 (may be NULL) SM cycles of the frame: association (affinity + assignment),
 *                update, initialisation, total -- the (asso_time, update_time, init_time) tuple
 *                tracking() returns (tracking/IterativeTracker.py:131,169-180); pam_sm_clock_khz converts
 *   d_out_vlist [S][T][R][PAM_VLIST_BYTES] u8 (may be NULL) per reported track: [0] usable views of this
 *                update (= length of its joints_views list, IterativeTracker.py:350), [1] length of the track's view
 *                list (= len(poses2d), ivclabpose.py:276), [2 ...] camera of every view-list entry in dict-insertion
 *                order, bit 7 set when that view was matched THIS frame (ivclabpose.py:277-283) */
#define PAM_VLIST_BYTES 10
int pam_track_sequences(pam_handle* h, void* d_state, int32_t S, int32_t T, int32_t frame0,
                        const float* d_dets, const int32_t* d_counts, int32_t* d_out_count,
                        int32_t* d_out_ids, float* d_out_joints, uint8_t* d_out_nviews,
                        int32_t* d_out_assoc, int32_t* d_out_timing, uint8_t* d_out_vlist, void* stream);

/* Per-sequence status words; synchronises `stream`.  h_status[s] (may be NULL) = hard error code
 * (0 = ok) | PAM_WARN_* bits << 8.  The reference has no capacity limits; here a frame that needs more
 * track slots / hypotheses / detections than configured DROPS the excess (the new track is not created,
 * the detection is ignored), tracking continues, and the sequence is flagged.  Returns PAM_E_CAPACITY if
 * any sequence carries a warning, PAM_E_INTERNAL for a hard error, else PAM_OK. */
#define PAM_WARN_TRACKS 1
#define PAM_WARN_HYPOTHESES 2
#define PAM_WARN_DETECTIONS 4
int pam_track_status(pam_handle* h, const void* d_state, int32_t S, int32_t* h_status, void* stream);

/* Decision margins of a run (libraries built with -DPAM_MARGIN only, else PAM_E_INVALID): for every
 * sequence n_margins doubles = the smallest distance any decision came to flipping since the last reset:
 * |c| of "c > 0" (IterativeTracker.py:143), |A| of "A < 0" (matching.py:248), |ra-rb|/max of the ray rule
 * (matching.py:272), |believe - conf_threshold| (IterativeTracker.py:59), init-mode |A| (f32) and row-sum
 * difference (matching.py:287-294), |cost - 1| of the veto (hypothesis.py:66), smallest positive affinity. */
int pam_track_margins(pam_handle* h, const void* d_state, int32_t S, double* h_margins, void* stream);

/* Launch shape the tracker would use for S sequences: out[0] = warps per sequence, [1] = sequences per
 * CTA, [2] = threads per CTA, [3] = resident CTAs per SM, [4] = dynamic shared memory per CTA (bytes),
 * [5] = registers per thread, [6] = working-set bytes per sequence, [7] = detection buffers per sequence. */
int pam_track_launch_info(pam_handle* h, int32_t S, int32_t* out8);

/* SM clock (kHz) the cycle counts of d_out_timing refer to. */
int pam_sm_clock_khz(pam_handle* h, int32_t* khz);

/* Same as pam_track_sequences with HOST buffers: allocates/reuses device workspace inside the
 * handle, copies in, runs, copies out, synchronises.  `fresh` != 0 restarts the trackers first.
 * Capacity warnings do not fail the call (query them with pam_track_host_status). */
int pam_track_sequences_host(pam_handle* h, int32_t S, int32_t T, int32_t frame0, int32_t fresh,
                             const float* h_dets, const int32_t* h_counts, int32_t* h_out_count,
                             int32_t* h_out_ids, float* h_out_joints, uint8_t* h_out_nviews,
                             int32_t* h_out_assoc, int32_t* h_out_timing, uint8_t* h_out_vlist);

/* pam_track_status for the internal state of the _host path. */
int pam_track_host_status(pam_handle* h, int32_t S, int32_t* h_status);

/* ---- stream mode: the reference's call pattern, one tracking() call per frame (ivclabpose.py:257) ----------
 * ONE sequence.  pam_stream_open starts a resident kernel (one CTA) that keeps the tracker state on chip and polls
 * a command word in pinned, device-mapped host memory; pam_stream_buffers returns HOST pointers into that slot:
 * the caller writes the frame's detections (PACKED: cameras back to back, no padding) and counts there,
 * calls pam_stream_step(frame_id) -- which returns when the results are in the slot -- and reads count / ids /
 * joints / nviews / assoc / timing from it.  No launch, no copies, no state round trip per frame.  The kernel
 * leaves after about one idle second (the state is stored; the next step starts it again).  The tracker state is
 * the one of the _host path (S = 1): pam_track_state_to_host and pam_track_sequences_host close the stream first. */
typedef struct pam_stream_views {
    float* dets;            /* [sum of counts][J][3]   in: the detections of camera 0, 1, ... back to back
                               (room for V * D of them)                                                   */
    int32_t* counts;        /* [V]                     in  */
    int32_t* out_count;     /* [1]                     out */
    int32_t* out_ids;       /* [max_tracks]                */
    float* out_joints;      /* [max_tracks][J][3]          */
    uint8_t* out_nviews;    /* [max_tracks][J]             */
    int32_t* out_assoc;     /* [V][D]                      */
    int32_t* out_timing;    /* [4] SM cycles: association, update, initialisation, total */
    int32_t* out_status;    /* [1] hard error code | PAM_WARN bits << 8 */
    uint8_t* out_vlist;     /* [max_tracks][PAM_VLIST_BYTES] */
} pam_stream_views;
int pam_stream_open(pam_handle* h, int32_t fresh);
int pam_stream_buffers(pam_handle* h, pam_stream_views* out);
int pam_stream_step(pam_handle* h, int32_t frame_id);        /* = submit + wait */
int pam_stream_submit(pam_handle* h, int32_t frame_id);      /* returns at once; host work may overlap the frame */
int pam_stream_wait(pam_handle* h);
int pam_stream_close(pam_handle* h);

/* Copy the internal state of the _host path (S sequences) to a host buffer of S*seq_bytes. */
int pam_track_state_to_host(pam_handle* h, int32_t S, void* h_state);

/* ---- stateless batched ops (device pointers, asynchronous on `stream`) ------------------------
 * Camera constants come from pam_set_cameras; J = cfg.num_joints; up to 32 cameras. */

/* Camera ingest on the device (Camera.__init__ + GetCameraParameters, ivclabpose.py:35-46,162-181):
 * d_K [V][3][3], d_RT [V][3][4] f32 -> d_RKinv [V][9] f32, d_pos [V][3] f64, d_F [V][V][9] f32, the arrays
 * pam_set_cameras takes (P is an input of the reference, not derived).  No handle needed.  Float32
 * arithmetic in the reference's association order; equal to the host ingest up to float32 rounding
 * (torch/LAPACK round differently in the last bits), so parity runs use the host ingest. */
int pam_camera_ingest(int32_t device, int32_t V, const float* d_K, const float* d_RT, float* d_RKinv, double* d_pos,
                      float* d_F, void* stream);

/* Camera.projectPoints_parallel (ivclabpose.py:91-98) for all cameras at once:
 * d_points3d [n_points][3] f64 -> d_out_vu [V][n_points][2] f64 as (v, u). */
int pam_project_points(pam_handle* h, const double* d_points3d, int32_t n_points, double* d_out_vu, void* stream);

/* Association affinity of tracking() (tracking/IterativeTracker.py:137-149) for all cameras:
 * d_tracks3d [n][J][3] f64, d_dt [n] i32 (frame - last pose time), d_dets [V][max_dets][J][3] f64
 * (v,u,conf), d_counts [V] i32 -> d_aff [V][n][max_dets] f64 (0 beyond counts). */
int pam_assoc_affinity(pam_handle* h, const double* d_tracks3d, const int32_t* d_dt, const double* d_dets,
                       const int32_t* d_counts, int32_t n_tracks, int32_t max_dets, double* d_aff, void* stream);

/* scipy.optimize.linear_sum_assignment (call sites tracking/IterativeTracker.py:79,150), batched:
 * d_cost [batch][n_rows][n_cols] f64 -> d_col4row [batch][n_rows] i32 (-1 = unassigned);
 * n_rows, n_cols <= 64. */
int pam_assign(pam_handle* h, const double* d_cost, int32_t batch, int32_t n_rows, int32_t n_cols, int32_t maximize,
               int32_t* d_col4row, void* stream);

/* epipolar_affinity_parallel (utils/matching.py:115-151): d_pose [M][J][3] f64 (v,u,conf),
 * d_cam [M] i32 camera index per pose -> d_dist [M][M][J] f64, d_mean [M][M] f64. */
int pam_epipolar_pairs(pam_handle* h, const double* d_pose, const int32_t* d_cam, int32_t M, double* d_dist,
                       double* d_mean, void* stream);

/* epipolar_affinity (utils/matching.py:93-113), float32 stores: d_aff [M][M] f32 (25 between poses of
 * one camera, 0 on the diagonal), d_dist_or_null [M][M][J] f32. */
int pam_epipolar_allpairs(pam_handle* h, const double* d_pose, const int32_t* d_cam, int32_t M, float* d_aff,
                          float* d_dist_or_null, void* stream);

/* epipolar_distance (utils/matching.py:50-91) for `batch` pose pairs: d_pose1/2 [batch][J][3] f64,
 * d_cam1/2 [batch] i32 -> d_out [batch][J][2] f64 = [d(x1, F x2), d(x2, F^T x1)]. */
int pam_epipolar_distance(pam_handle* h, const double* d_pose1, const double* d_pose2, const int32_t* d_cam1,
                          const int32_t* d_cam2, int32_t batch, double* d_out, void* stream);

/* Greedy_matching (utils/matching.py:243-295) for `batch` joints over n_views views:
 * mode 0 'update': d_affinity_f64 [batch][n][n], d_uv [batch][n][2] (u,v), d_next [batch][3];
 * mode 1 'init':   d_affinity_f32 [batch][n][n];   d_cam [n] i32 -> d_keep [batch][n] u8. */
int pam_view_filter(pam_handle* h, int32_t mode, const double* d_affinity_f64, const float* d_affinity_f32,
                    const double* d_uv, const int32_t* d_cam, const double* d_next, int32_t batch, int32_t n_views,
                    uint8_t* d_keep, void* stream);

/* SVD_pose_kernel_jf / _parallel / SVD_pose_kernel (utils/construction.py:64-131): d_pose
 * [batch][n_views][J][3] f64, d_cam [batch][n_views] i32, d_weight [batch][n_views] f64
 * (exp(-lambda_t T)), d_keep_or_null [batch][J][n_views] u8, d_next_or_null [batch][J][3] f64 (value of
 * joints with < 2 views; NaN if null) -> d_out [batch][J][3] f64. */
int pam_triangulate(pam_handle* h, const double* d_pose, const int32_t* d_cam, const double* d_weight,
                    const uint8_t* d_keep_or_null, const double* d_next_or_null, int32_t batch, int32_t n_views,
                    double* d_out, void* stream);

/* back_project_ray (utils/matching.py:10-17) + line2point_distance_3D (utils/calculate.py:26-32):
 * d_uv [n][2] f64 (u,v) -> d_dirs_or_null [n][3] unit rays, d_dist_or_null [n] distance of
 * d_points3d_or_null [n][3] to the rays. */
int pam_ray_distance(pam_handle* h, int32_t camera, const double* d_uv, const double* d_points3d_or_null, int32_t n,
                     double* d_dist_or_null, double* d_dirs_or_null, void* stream);

/* get_believe (utils/calculate.py:8-14), batched: d_pose [batch][J][3] f64 -> d_out [batch] f64 mean
 * confidence over the joints with conf >= 0 (NaN when there is none). */
int pam_mean_confidence(pam_handle* h, const double* d_pose, int32_t batch, double* d_out, void* stream);

/* Hypothesis.calculate_cost (tracking/hypothesis.py:53-68) of one hypothesis -- d_hyp_pose
 * [n_views][J][3] f64, d_hyp_cam [n_views] i32 -- against `batch` candidate detections d_other_pose
 * [batch][J][3] f64 of camera other_cam -> d_cost [batch] f64, d_veto [batch] u8 (uses
 * cfg.epi_threshold and cfg.veto_believe). */
int pam_hypothesis_cost(pam_handle* h, const double* d_hyp_pose, const int32_t* d_hyp_cam, int32_t n_views,
                        const double* d_other_pose, int32_t other_cam, int32_t batch, double* d_cost, uint8_t* d_veto,
                        void* stream);

/* PCP / MPJPE counters of Evaluate3DPose_PCP (evalmodel.py:120-206, eval/transformation.py:5-39,
 * eval/numeric.py:5-25) straight from the tracker's output tensors: d_out_count [S][T], d_out_joints
 * [S][T][max_tracks][J][3] f32; d_gt [S][T][P][14][3] f64 (Shelf/Campus joint order), d_gt_valid
 * [S][T][P] u8; frames frame_begin <= t < frame_end; alpha = 0.5 in the reference.  Accumulates (does
 * not reset) d_counters [P][10][2] u64 = (correct, evaluated) per actor and part (9 limbs + hip-head)
 * and d_mpjpe [2] f64 = (sum of joint errors, joints) -- the counters a multi-GPU run all-reduces. */
int pam_eval_pcp(pam_handle* h, const int32_t* d_out_count, const float* d_out_joints, const double* d_gt,
                 const uint8_t* d_gt_valid, int32_t S, int32_t T, int32_t P, int32_t max_tracks, int32_t remap_coco17,
                 int32_t frame_begin, int32_t frame_end, double alpha, uint64_t* d_counters, double* d_mpjpe,
                 void* stream);

/* Matching step of EvaluatePanoptic.evaluate (evalmodel.py:291-320): for every predicted pose of every
 * frame the MPJPE (mm, visible joints) to its closest ground-truth body.  d_out_count / d_out_joints:
 * tracker outputs (J = 17 COCO or 19 COCO-19); d_gt_mm [S][T][max_gt][14][3] f64 in millimetres,
 * d_gt_vis [S][T][max_gt][14] u8, d_n_gt [S][T] i32 -> d_mpjpe [S][T][max_tracks] f64, d_gt_index
 * [S][T][max_tracks] i32 (-1: no ground truth in the frame / slot unused).  AP, recall and MPJPE are
 * then list reductions on the host (pam_b200.evaluate.panoptic_metrics). */
int pam_eval_panoptic_match(pam_handle* h, const int32_t* d_out_count, const float* d_out_joints, const double* d_gt_mm,
                            const uint8_t* d_gt_vis, const int32_t* d_n_gt, int32_t S, int32_t T, int32_t max_gt,
                            int32_t max_tracks, double* d_mpjpe, int32_t* d_gt_index, void* stream);

/* ---- alternatives the reference keeps in its API without a live caller ----------- */

/* One-Euro filter bank (tracking/OneEuroFilter.py:12-77): n channels sharing their time stamps, e.g. the 3 x J
 * coordinates of one track (tracking/IterativeTracker.py:231-237).  d_state [n][4] f64 = {last raw value, filtered value,
 * filtered derivative, initialised (0/1)}, zero = fresh; `freq` is kept by the caller like the reference keeps it
 * (1 / (t - t_last) once both time stamps are truthy).  d_out [n] = filtered values; bit-identical to the python class. */
int pam_one_euro(pam_handle* h, const double* d_x, int32_t n, double freq, double mincutoff, double beta, double dcutoff,
                 double* d_state, double* d_out, void* stream);

/* Kalman smoother bank (tracking/KalmanFilter.py:4-65: cv2.KalmanFilter(9, 3) per 3-D joint, constant acceleration,
 * float32).  d_state [n][90] f32 = state vector (9) + error covariance (81); initialise the first three state entries
 * with the joint and the rest with zeros (KalmanFilter.__init__).  One call = the reference's predict(pt3d): correct with
 * d_meas_or_null [n][3] f64 where d_has_meas_or_null [n] (null = all) is set, then predict; d_out [n][3] = predicted
 * positions.  Agrees with OpenCV to float32 rounding (it inverts the 3 x 3 innovation covariance by SVD). */
int pam_kalman9(pam_handle* h, const double* d_meas_or_null, const uint8_t* d_has_meas_or_null, int32_t n, double hz,
                float* d_state, double* d_out, void* stream);

/* top_down_pose_kernel (utils/construction.py:9-31): every camera pair triangulates all joints from its two views
 * (cv2.triangulatePoints' 4 x 4 system); the pair whose pose reprojects best into all cameras wins.
 * d_poses2d [batch][n_views][J][2] f64 (x, y), d_cam [batch][n_views] i32 -> d_pose3d [batch][J][3] f64,
 * d_pair_or_null [batch][2] i32 (the winning views), d_err_or_null [batch][n_views (n_views - 1) / 2] f64 (the summed
 * reprojection error of every pair, pairs in (0,1), (0,2), ... order). */
int pam_top_down(pam_handle* h, const double* d_poses2d, const int32_t* d_cam, int32_t batch, int32_t n_views,
                 double* d_pose3d, int32_t* d_pair_or_null, double* d_err_or_null, void* stream);

/* number of kernel launches issued through this handle so far (bench.py "gpu_launches") */
int64_t pam_launch_count(const pam_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* PAM_H */
