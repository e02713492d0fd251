"""Randomised shapes / parameters ON THE DEVICE against the oracle.

  python tools/gpu_fuzz.py --seeds 0 1000 [--large 0 150] [--max-tracks 32] [--log gpurun_out/fuzz_gpu.log]

Seeds are the configurations of tests/test_fuzz_hostemu.py (`_case`: 2-6 cameras, 1-5 people, 12-22 joints,
random thresholds / noise / misses / outliers / staggered entries / absences; `_large_case`: 6-8 cameras, 5-8
people, 23-32 joints).  The oracle side runs in a process pool on the host cores; every configuration is then
tracked by libpam.so on cuda:0 (pam_track_sequences, one launch per configuration) and compared: track ids,
reported sets, per-joint view counts and detection->track associations exact, joints within 0.5 mm / 1e-3."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def oracle_job(job):
    kind, seed = job
    from tests import test_fuzz_hostemu as tf
    from oracle import generic
    from pam_b200 import synth
    shape, params, kw, min_valid = (tf._case if kind == "case" else tf._large_case)(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    oo, oa, trk = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    return kind, seed, oo, oa, max((len(f[0]) for f in oo), default=0), len(trk.tracks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 1000])
    ap.add_argument("--large", type=int, nargs=2, default=[0, 0])
    ap.add_argument("--max-tracks", type=int, default=32)
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "fuzz_gpu.log"))
    a = ap.parse_args()
    import multiprocessing as mp
    import torch
    from tests import util, test_fuzz_hostemu as tf
    from pam_b200 import camera, synth, tracker
    jobs = [("case", s) for s in range(*a.seeds)] + [("large", s) for s in range(*a.large)]
    os.makedirs(os.path.dirname(a.log), exist_ok=True)
    log = open(a.log, "w")
    t0 = time.time()
    n_ok = n_bad = n_warn = 0
    worst = 0.0
    bad = []
    with mp.get_context("spawn").Pool(os.cpu_count() or 4) as pool:
        for kind, seed, oo, oa, max_rep, live in pool.imap_unordered(oracle_job, jobs, chunksize=4):
            shape, params, kw, min_valid = (tf._case if kind == "case" else tf._large_case)(seed)
            st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
            trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), params, 1, max_detections=st.dets.shape[2],
                                          max_tracks=a.max_tracks, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
            out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), nviews=True, assoc=True)
            status = int(trk.check(strict=False)[0])
            out = {k: v.cpu().numpy() for k, v in out.items()}
            trk.close()
            try:
                w = util.compare_with_oracle(out, 0, st, oo, oa)
                worst = max(worst, w)
                verdict = "ok" if status == 0 else f"ok-but-status-{status:#x}"
                n_ok += 1
                n_warn += status != 0
            except AssertionError as e:
                verdict = f"MISMATCH status={status:#x} {str(e)[:160]}"
                n_bad += 1
                bad.append((kind, seed))
            log.write(f"{kind} {seed:5d} V{shape.V} P{shape.P} J{shape.J} T{shape.T} live_tracks_end={live} "
                      f"max_reported={max_rep} {verdict}\n")
            log.flush()
    tail = (f"TOTAL {len(jobs)} configurations on {torch.cuda.get_device_name(0)}: {n_ok} identical to the oracle "
            f"({n_warn} of them with a capacity warning), {n_bad} mismatches {bad}; max |dX| {worst:.3e} m; "
            f"{time.time() - t0:.0f} s; max_tracks={a.max_tracks}")
    log.write(tail + "\n")
    log.close()
    print(tail)
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
