// pam_host.h -- host-side helpers shared by the C ABI (pam_lib.cu) and the test-only host
// harness (tests/hostemu/hostemu.cpp): pam_config validation and the launch-constant DevCfg.
#pragma once
#include <string>
#include "../../include/pam.h"
#include "pam_track.h"

namespace pam {

// true when the configuration fits the stateful tracker's compile-time capacities (<= 8 cameras);
// rigs with up to 32 cameras are served by the stateless batched ops only.
inline bool tracker_capable(const pam_config& p) { return p.num_cameras <= PAM_MAX_V; }

// nbuf / raw_in_arena: the working-set flavour (arena_layout); the global state layout does not depend on them
inline int make_devcfg(const pam_config& p, DevCfg& c, std::string& err, int nbuf = 2, bool raw_in_arena = true) {
    auto bad = [&](const char* m) { err = m; return (int)PAM_E_INVALID; };
    if (p.num_cameras < 1 || p.num_cameras > 32) return bad("num_cameras outside 1..32");
    if (p.num_joints < 1 || p.num_joints > PAM_MAX_J) return bad("num_joints outside 1..32");
    if (p.max_detections < 1 || p.max_detections > PAM_MAX_D) return bad("max_detections outside 1..16");
    if (p.max_tracks < 1 || p.max_tracks > PAM_MAX_TRK) return bad("max_tracks outside 1..32");
    if (p.max_age < 0 || p.max_age + 2 > PAM_HIST) return bad("max_age + 2 exceeds the history ring (12)");
    if (p.stale_window < 0 || p.stale_window + 1 > PAM_MAX_AGEW) return bad("stale_window outside 0..7");
    if (!(p.sigma > 0.0) || !(p.arm_sigma > 0.0)) return bad("sigma / arm_sigma must be positive");
    c = DevCfg();
    c.V = p.num_cameras; c.J = p.num_joints; c.D = p.max_detections; c.max_trk = p.max_tracks;
    c.max_rep = (p.max_report == 0 || (int)p.max_report > p.max_tracks) ? p.max_tracks : (int)p.max_report;
    c.max_hyp = p.num_cameras * p.max_detections < PAM_MAX_HYP ? p.num_cameras * p.max_detections : PAM_MAX_HYP;
    c.n_init = p.n_init; c.max_age = p.max_age; c.min_valid = p.min_valid_joints; c.stale_window = p.stale_window;
    c.arm_mask = p.arm_joint_mask;
    c.rad[0] = gaussian_weights(p.sigma, c.gw[0]);
    c.rad[1] = gaussian_weights(p.arm_sigma, c.gw[1]);
    if (c.rad[0] < 0 || c.rad[1] < 0) return bad("sigma too large: Gaussian radius int(4 sigma + .5) > 8");
    for (int T = 0; T < PAM_MAX_AGEW; ++T) c.w_age[T] = exp(-p.lambda_t * (double)T);
    c.inv_J = 1.0f / (float)c.J; c.inv_D = 1.0f / (float)c.D; c.inv_V = 1.0f / (float)c.V; c.inv_VD = 1.0f / (float)(c.V * c.D);
    c.inv_joint_thr = 1.0 / p.joint_threshold;
    c.joint_thr2 = p.joint_threshold * p.joint_threshold;
    for (int dt = 0; dt < 16; ++dt) {
        c.inv_denom_tab[dt] = 1.0 / (p.alpha2d * (double)dt);
        c.inv_decay_tab[dt] = 1.0 / exp(p.lambda_a * (double)dt);
    }
    c.conf_thr = p.conf_threshold; c.epi_thr = p.epi_threshold; c.joint_thr = p.joint_threshold;
    c.alpha2d = p.alpha2d; c.lambda_a = p.lambda_a; c.veto_believe = p.veto_believe;
    c.fail_limit = (double)p.num_joints / 3.0;
    c.init_thr_f32 = (float)p.init_threshold;
    // two-pass affinity: probe J - min_valid joints (at least 3): a pair without enough hits by then is hopeless
    c.aff_probe = 0;
    if (c.J >= 8 && c.min_valid >= 0 && c.min_valid < c.J) {
        int ja = c.J - c.min_valid;
        if (ja < 3) ja = 3;
        if (ja <= c.J - 2) c.aff_probe = ja;
    }
    state_layout(c);
    if (tracker_capable(p)) arena_layout(c, nbuf, raw_in_arena);
    return PAM_OK;
}

inline void fill_layout(const DevCfg& c, pam_state_layout& L) {
    L.seq_bytes = c.seq_bytes; L.off_header = c.off_hdr; L.off_meta = c.off_meta; L.off_hist = c.off_hist;
    L.off_view = c.off_view; L.off_vel = c.off_vel; L.off_nviews = c.off_nv; L.off_margin = c.off_margin;
    L.meta_ints = (int32_t)(sizeof(TrkMeta) / 4); L.hist_ring = PAM_HIST; L.max_views = PAM_MAX_V;
    L.max_order = PAM_MAX_TRK; L.header_ints = 8; L.n_margins = MG_COUNT;
}

}  // namespace pam
