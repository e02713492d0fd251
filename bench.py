#!/usr/bin/env python
"""bench.py -- synthetic multi-view frames/s (triangulated + associated) of the part-aware tracker.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], batched as configs[4]): Shelf-shaped synthetic streams -- 5
cameras, 4 people, 14 joints, 3200 frames per sequence -- S independent sequences per GPU (weak
scaling: every rank tracks its own S sequences; sequences never span GPUs).  One "step" = one pass
of the tracker over all S x T frames of the rank, starting from empty trackers.
`--scaling strong --sequences-total 1024` is configs[4] as written (a fixed total sharded over the
ranks, NCCL all-gather of the result blocks inside the step); the default weak run also reports it
as `config5_strong_1024`.

Printed JSON keys follow the driver contract:
  value      whole-job frames/s with detections already resident in HBM (restart + kernel), CUDA
             events on the launching stream, max over ranks
  e2e        the same through the reference-facing C-ABI call pam_track_sequences_host with pinned
             HOST buffers: H2D of the detections, kernel, D2H of ids/joints, inside the timed region
  roofline   the tracker kernel against the measured HBM copy peak (MEASURED_PEAKS.json); the
             kernel is FP64-latency/issue bound, so the fraction is small by construction
             (DESIGN.md section "Roofline")
  cpu_baseline  the numpy oracle (oracle/generic.py, a bit-identical restatement of the reference's
             own per-frame path) timed on one host core over a bounded prefix of the same workload
  parity_sample   K sequences of the TIMED batch compared with the oracle over all frames (checker leg)
  other_configs   Campus / Panoptic / Dense configurations of BASELINE.json through the same library
  per_frame_api   the literal drop-in call pattern, one IterativeTracker.tracking() call per frame
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SHAPE = "shelf"
METRIC = "synthetic multi-view frames/sec (triangulated+associated)"


def workload_name(shape_name, sh, T, S):
    """The workload both arms are quoted on (config.workload)."""
    return (f"{shape_name}: {sh.V} cameras x {sh.P} people x {sh.J} joints x {T} frames, {S} independent "
            f"sequences per GPU")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sequences", type=int, default=2368,
                    help="independent sequences per GPU (148 SMs x 16: one warp per sequence, 16 resident per SM; 36 GB of "
                         "pinned host memory per rank)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --sequences-total sequences sharded over the ranks + NCCL all-gather of the results (config 5)")
    ap.add_argument("--sequences-total", type=int, default=1024)
    ap.add_argument("--parity-sequences", type=int, default=4, help="sequences of the timed batch checked against the oracle")
    ap.add_argument("--no-extras", action="store_true", help="skip parity sample, config 5, other configurations, per-frame API")
    ap.add_argument("--frames", type=int, default=None, help="frames per sequence (default: the shape's 3200)")
    ap.add_argument("--shape", default=SHAPE)
    ap.add_argument("--cpu-frames", type=int, default=4500, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# data
# ----------------------------------------------------------------------------------------------
def generate(shape, seq_ids, T, dets_out, counts_out, gt_out=None, workers=None):
    from concurrent.futures import ThreadPoolExecutor
    import pam_b200  # noqa: F401
    from pam_b200 import synth
    rig = synth.make_rig(shape)

    def one(k):
        st = synth.make_stream(shape, seq_ids[k], T, rig=rig)
        dets_out[k] = st.dets
        counts_out[k] = st.counts
        if gt_out is not None:
            gt_out[k] = st.gt

    with ThreadPoolExecutor(workers or min(32, os.cpu_count() or 8)) as ex:
        list(ex.map(one, range(len(seq_ids))))
    return rig


def bind_to_gpu_numa_node(device_index):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal CPU affinity) BEFORE the pinned host
    buffers are allocated, so that first-touch places them on the GPU's NUMA node: on a two-socket 8-GPU
    box the H2D/D2H legs of `e2e` otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled during the timed region (B200_PROFILING.md recipe: the
    nvidia-smi query below; read through NVML directly when pynvml is importable, so that a sub-second
    timed region still gets tens of samples)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_evt = index, [], threading.Event()

    def _nvml_row(self):
        """Same fields through NVML (what nvidia-smi reads), cheap enough to sample every 20 ms."""
        import pynvml as nv
        if not hasattr(self, "_h"):
            nv.nvmlInit()
            self._h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self._max = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self._h))
        act = lambda bit: "Active" if r & bit else "Not Active"
        # bits: sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40 (nvml.h)
        return [str(sm), str(self._max), "", act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        use_nvml = True
        while not self.stop_evt.is_set():
            if use_nvml:
                try:
                    self.rows.append(self._nvml_row())
                    self.stop_evt.wait(0.02)
                    continue
                except Exception:
                    use_nvml = False
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in o.strip().split(",")]
                if len(p) >= 7:
                    self.rows.append(p)
            except Exception:
                pass
            self.stop_evt.wait(0.15)

    def summary(self):
        self.stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's per-frame path (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    shape, seq_id, frames, warm = args
    import pam_b200  # noqa: F401
    from pam_b200 import synth
    from oracle import generic
    st = synth.make_stream(shape, seq_id, frames + warm)
    cams = generic.build_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    trk = generic.Tracker(synth.tracker_params(shape), st.shape.arm_joints, 10)
    V = st.shape.V
    inputs = [(st.frame_boxes(t), st.frame_detections(t)) for t in range(st.T)]
    for t in range(warm):        # first frames discarded like src/testmodel.py:86
        trk.tracking(t, cams, [None] * V, inputs[t][0], inputs[t][1], "SVD")
    t0 = time.perf_counter()
    for t in range(warm, st.T):  # timer around tracking() only, like src/evalmodel.py:79-82
        trk.tracking(t, cams, [None] * V, inputs[t][0], inputs[t][1], "SVD")
    return time.perf_counter() - t0


def cpu_sample(shape, frames, procs, seq0=900000, warm=10):
    """frames/s of the oracle over `procs` processes, one sequence prefix each."""
    if procs == 1:
        el = _cpu_worker((shape, seq0, frames, warm))
        return frames / el, el
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        t0 = time.perf_counter()
        els = pool.map(_cpu_worker, [(shape, seq0 + k, frames, warm) for k in range(procs)])
        wall = time.perf_counter() - t0
    # aggregate rate = sum of the per-process rates (each process times only tracking())
    return sum(frames / e for e in els), wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pam_b200 import synth
    sh = synth.SHAPES[a.shape]
    cores = os.cpu_count() or 1
    frames = 250                     # per process per step: ~2 s of CPU work
    rates = []
    for it in range(a.warmup + a.steps):
        r, _ = cpu_sample(a.shape, frames, cores, seq0=900000 + 1000 * it)
        if it >= a.warmup:
            rates.append(r)
    val = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * frames * cores / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.shape, sh, a.frames or sh.T, a.sequences),
                   "sample": f"each step: {cores} sequences x {frames} frames of that workload, one process per host core"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {frames} frames, numpy oracle (bit-identical restatement of "
                                   "the reference's IterativeTracker.tracking), timer around tracking() only"},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# checker leg: a sample of the TIMED batch against the oracle
# ----------------------------------------------------------------------------------------------
def _oracle_worker(args):
    shape, seq_id, T = args
    import pam_b200  # noqa: F401
    from pam_b200 import synth
    from oracle import generic
    st = synth.make_stream(shape, seq_id, T)
    oo, oa, _ = generic.run_stream(st, synth.tracker_params(shape), st.shape.arm_joints, 10, trace=True)
    return seq_id, oo, oa


def parity_sample(shape, seq_ids, local_index, T, big_out, rerun):
    """Compare K sequences of the timed batch with oracle.generic.run_stream (ids, counts, per-joint view counts,
    detection->track associations exact; joints within 0.5 mm / 1e-3).  `big_out`: count/ids/joints of those
    sequences pulled from the timed batch; `rerun`: the same sequences tracked again as a small batch with the
    optional outputs (view counts, associations) switched on."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(len(seq_ids)) as pool:
        res = pool.map(_oracle_worker, [(shape, sid, T) for sid in seq_ids])
    worst, frames, reports = 0.0, 0, 0
    for k, (sid, oo, oa) in enumerate(res):
        for t in range(T):
            ids, joints, views = oo[t]
            n = int(big_out["count"][k, t])
            assert n == len(ids) == int(rerun["count"][k, t]), f"sequence {sid} frame {t}: {n} tracks, oracle {len(ids)}"
            r = min(n, big_out["ids"].shape[2])            # rows the output stride keeps (max_report)
            ids, joints, views = ids[:r], joints[:r], views[:r]
            assert np.array_equal(big_out["ids"][k, t, :r], ids) and np.array_equal(rerun["ids"][k, t, :r], ids), (sid, t)
            if r:
                got = big_out["joints"][k, t, :r].astype(np.float64)
                assert np.array_equal(big_out["joints"][k, t, :r], rerun["joints"][k, t, :r]), (sid, t)
                err = np.abs(got - joints)
                assert np.all(err <= np.maximum(5e-4, 1e-3 * np.abs(joints))), f"sequence {sid} frame {t}: {err.max():.3e} m"
                worst = max(worst, float(err.max()))
                assert np.array_equal(rerun["nviews"][k, t, :r], views), (sid, t)
            for c, a in enumerate(oa[t]):
                assert np.array_equal(rerun["assoc"][k, t, c, :len(a)], a), (sid, t, c)
            reports += n
        frames += T
    return {"sequences": [int(x) for x in seq_ids], "positions_in_batch": [int(x) for x in local_index], "frames": frames,
            "track_reports": reports, "max_dx_m": worst, "ids_counts_nviews_assoc": "identical",
            "checker": "oracle.generic.run_stream (numpy restatement of the reference path)"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def frame_bytes(V, D, J, n_out):
    """Algorithmic bytes per frame, SURVEY.md section 8d: detections in + joints/ids out."""
    return 4.0 * (3 * V * D * J + 3 * n_out * J + n_out)


def time_kernel(trk, dets, counts, out, reps=3):
    import torch
    best = 1e30
    for it in range(reps + 1):
        trk.restart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trk.run(dets, counts, out=out, frame0=0)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    return best


def tiled_batch(shape, unique, S, T, dev):
    """S sequences on the device as a tiling of `unique` seeded ones (side configurations only)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from pam_b200 import synth
    rig = synth.make_rig(shape)
    with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as ex:
        streams = list(ex.map(lambda s: synth.make_stream(shape, 500000 + s, T, rig=rig), range(unique)))
    du = torch.from_numpy(np.stack([s.dets for s in streams])).to(dev)
    cu = torch.from_numpy(np.stack([s.counts for s in streams])).to(dev)
    reps = (S + unique - 1) // unique
    return rig, du.repeat(reps, 1, 1, 1, 1, 1)[:S].contiguous(), cu.repeat(reps, 1, 1)[:S].contiguous()


def other_config(shape, S, unique, dev, peak, cpu_frames):
    """One of the other BASELINE.json configurations through the same kernel: batched + single stream."""
    import torch
    from pam_b200 import camera, synth, tracker
    sh = synth.SHAPES[shape]
    T = sh.T
    rig, dets, counts = tiled_batch(shape, unique, S, T, dev)
    MT = 8 if sh.P <= 6 else 12
    cams = camera.GetCameraParameters(rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(shape), S, max_detections=sh.P, max_tracks=MT,
                                  arm_joints=sh.arm_joints, device=dev.index or 0)
    out = trk.alloc_outputs(T, nviews=False, assoc=False)
    ms = time_kernel(trk, dets, counts, out)
    trk.check(strict=False)
    n_out = float(out["count"].sum().item()) / (S * T)
    info = trk.launch_info()
    b = frame_bytes(sh.V, sh.P, sh.J, n_out)
    t1 = tracker.SequenceTracker(cams, synth.tracker_params(shape), 1, max_detections=sh.P, max_tracks=MT,
                                 arm_joints=sh.arm_joints, device=dev.index or 0)
    o1 = t1.alloc_outputs(T, nviews=False, assoc=False)
    ms1 = time_kernel(t1, dets[:1].contiguous(), counts[:1].contiguous(), o1)
    rate, el = cpu_sample(shape, cpu_frames, 1)
    res = {"workload": f"{shape}: {sh.V} cameras x {sh.P} people x {sh.J} joints x {T} frames, {S} sequences "
                       f"({unique} seeded ones tiled) on one GPU", "value": S * T / (ms * 1e-3), "unit": "frames/s",
           "kernel_ms": ms, "launch": info,
           "roofline": {"bound": "hbm", "achieved": b * S * T / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": b * S * T / (ms * 1e-3) / 1e9 / peak, "bytes_per_frame": b, "traffic": None},
           "single_stream": {"value": T / (ms1 * 1e-3), "unit": "frames/s", "us_per_frame": ms1 * 1e3 / T},
           "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                            "sample": f"first {cpu_frames} frames of one {shape} sequence ({el:.1f} s)"}}
    trk.close(); t1.close()
    del dets, counts, out, o1
    torch.cuda.empty_cache()
    return res


def dense_config(dev, peak):
    """BASELINE.json configs[3]: one dense frame (31 cameras x 64 people x 19 joints, M = 1984) through the stateless
    ops, L2 flushed between timed iterations (tools/dense_bench.py), plus a bounded oracle sample."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dense_bench", os.path.join(ROOT, "tools", "dense_bench.py"))
    mod = importlib.util.module_from_spec(spec)
    import contextlib, io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        spec.loader.exec_module(mod)
    ops_rows = [{"op": name, "ms": ms, "algorithmic_bytes": b,
                 "roofline": {"bound": "hbm", "achieved": b / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": b / ms / 1e6 / peak}}
                for name, ms, b, fl in mod.rows]
    frame_ms = sum(r["ms"] for r in ops_rows if "per-joint" not in r["op"])
    # bounded CPU sample: 300 of the M(M-1)/2 pose pairs through the oracle's epipolar_distance, scaled
    from oracle import generic
    from pam_b200 import synth
    st = mod.st
    ocams = generic.build_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    rng = np.random.default_rng(0)
    M = mod.M
    t0 = time.perf_counter()
    n = 0
    while n < 300:
        a, b = (int(x) for x in rng.integers(0, M, 2))
        if mod.cam_idx[a] == mod.cam_idx[b]:
            continue
        generic.epipolar_distance(ocams[mod.cam_idx[a]], mod.poses[a], ocams[mod.cam_idx[b]], mod.poses[b])
        n += 1
    per_pair = (time.perf_counter() - t0) / n
    pairs = M * (M - 1) // 2 - mod.V * (mod.P * (mod.P - 1) // 2)
    return {"workload": f"dense crowd frame: {mod.V} cameras x {mod.P} people x {mod.J} joints, M = {M} detections "
                        "(all-pairs affinity + triangulation + association + assignment)",
            "value": 1e3 / frame_ms, "unit": "frames/s", "frame_ms": frame_ms, "ops": ops_rows,
            "cpu_baseline": {"value": 1.0 / (per_pair * pairs), "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": f"300 of the {pairs} cross-camera pose pairs of one frame through "
                                       f"oracle.generic.epipolar_distance ({per_pair * 1e6:.0f} us per pair), scaled to the "
                                       "all-pairs affinity alone"}}


def per_frame_api(dev):
    """The literal drop-in call pattern: IterativeTracker.tracking() once per frame (src/ivclabpose.py:257)."""
    from tests import util
    from pam_b200 import camera, synth, tracker
    D = util.load_dropin()
    st = synth.make_stream("shelf", 0, 2200)
    cams = camera.GetCameraParameters(st.rig)
    D.IterativeTracker.IterativeTracker.ARM_JOINTS = st.shape.arm_joints
    from types import SimpleNamespace
    trk = D.IterativeTracker.IterativeTracker(SimpleNamespace(**synth.tracker_params("shelf")))
    inputs = [(st.frame_boxes(t), st.frame_detections(t)) for t in range(st.T)]
    warm = 200
    for t in range(warm):
        trk.tracking(t, cams, [None] * 5, inputs[t][0], inputs[t][1], "SVD")
    t0 = time.perf_counter()
    for t in range(warm, st.T):
        trk.tracking(t, cams, [None] * 5, inputs[t][0], inputs[t][1], "SVD")
    el = (time.perf_counter() - t0) / (st.T - warm)
    res = {"workload": "one shelf stream, one IterativeTracker.tracking() call per frame (detections as python lists of "
                       "float64 arrays, like the reference's caller)", "value": 1.0 / el, "unit": "calls/s",
           "us_per_call": el * 1e6, "reported_ids_last_frame": [int(x) for x in trk.last_ids],
           "capacities": {"MAX_TRACKS": trk.MAX_TRACKS, "MAX_DETECTIONS": trk.MAX_DETECTIONS},
           "reference_core_ratio": None}
    # the same through the bare library: stream mode (resident kernel) and one launch per call
    raw = tracker.SequenceTracker(cams, synth.tracker_params("shelf"), 1, max_detections=4, max_tracks=8,
                                  arm_joints=st.shape.arm_joints, device=dev.index or 0)
    fs = raw.open_stream(True)
    for t in range(warm):
        fs.set_frame(inputs[t][1]); fs.step(t)
    t0 = time.perf_counter()
    for t in range(warm, st.T):
        fs.set_frame(inputs[t][1]); fs.step(t)
    el1 = (time.perf_counter() - t0) / (st.T - warm)
    cyc = int(fs.timing[3])
    res["stream_mode"] = {"value": 1.0 / el1, "unit": "calls/s", "us_per_call": el1 * 1e6, "device_us_per_frame": cyc / (raw.sm_clock_khz() * 1e-3),
                          "call": "pam_stream_step on the mapped slot (max_detections 4, max_tracks 8), packing by numpy.concatenate"}
    fr = [(np.ascontiguousarray(st.dets[None, t:t + 1]), np.ascontiguousarray(st.counts[None, t:t + 1])) for t in range(st.T)]
    out = None
    for t in range(warm):
        out = raw.run_host(fr[t][0], fr[t][1], frame0=t, fresh=(t == 0), nviews=False, out=out)
    t0 = time.perf_counter()
    for t in range(warm, st.T):
        out = raw.run_host(fr[t][0], fr[t][1], frame0=t, nviews=False, out=out)
    el2 = (time.perf_counter() - t0) / (st.T - warm)
    res["launch_per_call"] = {"value": 1.0 / el2, "unit": "calls/s", "us_per_call": el2 * 1e6,
                              "call": "pam_track_sequences_host, S = T = 1, numpy buffers (round-1 path)"}
    raw.close()
    return res


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import pam_b200  # noqa: F401
    from pam_b200 import camera, dist as pdist, synth, tracker

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank, local_rank, world = pdist.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    sh = synth.SHAPES[a.shape]
    T = a.frames or sh.T
    V, J, D = sh.V, sh.J, sh.P
    MT = 8 if sh.P <= 6 else 12
    strong = a.scaling == "strong"
    if strong:          # BASELINE.json configs[4] as written: a fixed total, sharded over the ranks
        my_ids = pdist.shard_sequences(a.sequences_total, rank, world)
        S = len(my_ids)
    else:
        S = a.sequences
        # host-memory guard: every rank pins its detections and result buffers; never take more than half of
        # this rank's share of the free host memory (an 8-rank run must not drive the box out of memory)
        try:
            import psutil
            per_seq = T * (V * D * J * 3 * 4 + V * 4) + T * ((sh.P + 1) * (J * 3 * 4 + 4) + 4)
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
            budget = 0.5 * psutil.virtual_memory().available / max(1, local_world)
            if S * per_seq > budget:
                S = max(148, int(budget / per_seq) // 148 * 148)
        except Exception:
            pass
        my_ids = [rank * S + s for s in range(S)]         # distinct seeds on every rank

    # ---- synthetic input, generated straight into pinned host memory; the ground truth goes to the device in
    #      chunks (it is only needed there) --------------------------------------------------------------
    h_dets = torch.empty((S, T, V, D, J, 3), dtype=torch.float32, pin_memory=True)
    h_counts = torch.empty((S, T, V), dtype=torch.int32, pin_memory=True)
    do_eval = (J == 14)                                 # PCP counters need the 14 Shelf/Campus joints
    d_gt = torch.empty((S, T, sh.P, J, 3), dtype=torch.float64, device=dev) if do_eval else None
    t0 = time.time()
    rig = None
    CH = 148
    for c0 in range(0, S, CH):
        c1 = min(S, c0 + CH)
        gt_chunk = np.empty((c1 - c0, T, sh.P, J, 3), np.float64) if do_eval else None
        rig = generate(a.shape, my_ids[c0:c1], T, h_dets.numpy()[c0:c1], h_counts.numpy()[c0:c1], gt_chunk)
        if do_eval:
            d_gt[c0:c1].copy_(torch.from_numpy(gt_chunk))
    gen_s = time.time() - t0
    cams = camera.GetCameraParameters(rig)
    # output rows per frame: one more than the people in the scene (8 track slots stay available for ghosts and
    # hand-overs; `count` still tells the true number if a frame ever reports more) -- 37 % less D2H traffic
    MR = sh.P + 1
    trk = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), S, max_detections=D, max_tracks=MT,
                                  arm_joints=sh.arm_joints, device=local_rank, max_report=MR)
    d_dets = h_dets.to(dev, non_blocking=True)
    d_counts = h_counts.to(dev, non_blocking=True)
    out = trk.alloc_outputs(T, nviews=False, assoc=False)
    # run counters, summed over ranks: [reports, frames, PCP correct, PCP evaluated, MPJPE sum (um), joints counted]
    counters = torch.zeros(6, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    from pam_b200 import evaluate
    pcp = torch.zeros((sh.P, 10, 2), dtype=torch.int64, device=dev)
    mpj = torch.zeros(2, dtype=torch.float64, device=dev)
    # NCCL result gather (strong scaling = config 5): fixed-stride result blocks of every rank, device resident
    gathered = None
    if strong and world > 1:
        S_max = (a.sequences_total + world - 1) // world
        assert S == S_max or a.sequences_total % world, "ragged shards are padded"
        gathered = {k: torch.empty((world,) + tuple(out[k].shape), dtype=out[k].dtype, device=dev) for k in ("count", "ids", "joints")}
    gather_ev = []

    def step(timed_events=None):
        trk.restart()                                    # empty trackers: every step does the same work
        if timed_events is not None:
            timed_events[0].record(stream)
        trk.run(d_dets, d_counts, out=out, frame0=0)
        if timed_events is not None:
            timed_events[1].record(stream)
        counters[0] = out["count"].sum()
        counters[1] = S * T
        if do_eval:                                      # PCP / MPJPE counters on device (evalmodel.py:120-206)
            pcp.zero_(); mpj.zero_()
            evaluate.pcp_counters(trk, out, d_gt, counters=pcp, mpjpe=mpj, frame_begin=3)
            counters[2] = pcp[:, :, 0].sum()
            counters[3] = pcp[:, :, 1].sum()
            counters[4] = (mpj[0] * 1e6).to(torch.int64)
            counters[5] = mpj[1].to(torch.int64)
        if timed_events is not None:
            timed_events[2].record(stream)
        pdist.reduce_counters(counters)
        if gathered is not None:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            pdist.all_gather_results(out, gathered)
            g1.record(stream)
            if timed_events is not None:
                gather_ev.append((g0, g1))

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize(dev)
    trk.check()
    launches0 = trk.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    kev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(a.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pdist.barrier()
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for k in range(a.steps):
        step(kev[k])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    pdist.barrier()
    elapsed_ms = pdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.summary()
    launches = trk.launches - launches0              # tracker + evaluator kernels of this handle
    trk.check()
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y, _ in kev]))
    eval_ms = float(np.mean([y.elapsed_time(z) for _, y, z in kev]))      # report count + PCP / MPJPE counters
    gather_ms = pdist.max_over_ranks(float(np.mean([x.elapsed_time(y) for x, y in gather_ev])), dev) if gather_ev else None
    reports = int(out["count"].sum().item())
    total_frames = (a.sequences_total if strong else world * S) * T
    value = total_frames * a.steps / (elapsed_ms * 1e-3)
    launch_info = trk.launch_info()

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    e2e = None
    if not a.no_e2e:
        ho = dict(count=torch.empty((S, T), dtype=torch.int32, pin_memory=True).numpy(),
                  ids=torch.empty((S, T, MR), dtype=torch.int32, pin_memory=True).numpy(),
                  joints=torch.empty((S, T, MR, J, 3), dtype=torch.float32, pin_memory=True).numpy(),
                  nviews=None, assoc=None)
        hd, hc = h_dets.numpy(), h_counts.numpy()
        for _ in range(max(1, min(a.warmup, 2))):
            trk.run_host(hd, hc, fresh=True, nviews=False, out=ho)
        pdist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            trk.run_host(hd, hc, fresh=True, nviews=False, out=ho)
        torch.cuda.synchronize(dev)
        el = pdist.max_over_ranks(time.perf_counter() - t0, dev)
        assert int(ho["count"].sum()) == reports, "host path and device path disagree"
        e2e = {"value": total_frames * a.steps / el, "unit": "frames/s",
               "h2d_bytes_per_step": int(h_dets.numel() * 4 + h_counts.numel() * 4) * world,
               "d2h_bytes_per_step": int(ho["count"].nbytes + ho["ids"].nbytes + ho["joints"].nbytes) * world,
               "ms_per_step": 1e3 * el / a.steps, "timer": "host wall clock around the synchronous C-ABI call"}
        del ho

    # ---- config 5 beside the weak run: 1024 sequences in total over the ranks, NCCL all-gather of the results ----
    config5 = None
    if not strong and not a.no_extras and a.shape == "shelf" and S * world >= 1024 and 1024 % world == 0:
        S5 = 1024 // world
        t5 = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), S5, max_detections=D, max_tracks=MT,
                                     arm_joints=sh.arm_joints, device=local_rank, max_report=MR)
        d5, c5 = d_dets[:S5], d_counts[:S5]
        o5 = t5.alloc_outputs(T, nviews=False, assoc=False)
        g5 = {k: torch.empty((world,) + tuple(o5[k].shape), dtype=o5[k].dtype, device=dev) for k in ("count", "ids", "joints")} if world > 1 else None
        ks, gs = [], []
        for it in range(3 + 5):
            t5.restart()
            pdist.barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record(stream)
            t5.run(d5, c5, out=o5, frame0=0)
            ev[1].record(stream)
            if g5 is not None:
                pdist.all_gather_results(o5, g5)
            ev[2].record(stream)
            torch.cuda.synchronize(dev)
            if it >= 3:
                ks.append(ev[0].elapsed_time(ev[1])); gs.append(ev[1].elapsed_time(ev[2]))
        k_ms = pdist.max_over_ranks(float(np.mean(ks)), dev)
        g_ms = pdist.max_over_ranks(float(np.mean(gs)), dev)
        gb = sum(o5[k].numel() * o5[k].element_size() for k in ("count", "ids", "joints")) * world
        config5 = {"workload": f"1024 independent shelf sequences x {T} frames in total, {S5} per GPU, NCCL all-gather of the "
                               "fixed-stride result blocks (count, ids, joints) to every rank",
                   "scaling": "strong", "value": 1024 * T / ((k_ms + g_ms) * 1e-3), "unit": "frames/s",
                   "kernel_ms": k_ms, "gather_ms": g_ms if world > 1 else 0.0, "gather_bytes_per_rank": gb if world > 1 else 0,
                   "value_without_gather": 1024 * T / (k_ms * 1e-3), "launch": t5.launch_info(),
                   "limiter": "frame-serial chain of one sequence (about 10-14 us per frame per thread group) once a GPU "
                              "holds fewer sequences than it has SM slots; the gather is NVLink traffic of a few GB"}
        t5.close()
        del o5, g5

    if rank != 0:
        return

    # ---- checker: K sequences of the timed batch against the oracle ----------------------------------
    parity = None
    if not a.no_extras and a.parity_sequences > 0:
        try:
            rng = np.random.default_rng(12345)
            pick = sorted(int(x) for x in rng.choice(S, size=min(a.parity_sequences, S), replace=False))
            big = {k: out[k][pick].cpu().numpy() for k in ("count", "ids", "joints")}
            tk = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), len(pick), max_detections=D, max_tracks=MT,
                                         arm_joints=sh.arm_joints, device=local_rank, max_report=MR)
            rr = tk.run(d_dets[pick].contiguous(), d_counts[pick].contiguous(), nviews=True, assoc=True)
            tk.check()
            rr = {k: v.cpu().numpy() for k, v in rr.items()}
            tk.close()
            parity = parity_sample(a.shape, [my_ids[k] for k in pick], pick, T, big, rr)
        except AssertionError as e:
            parity = {"error": "MISMATCH: " + str(e)[:300]}

    # ---- BASELINE.json configs[1] taken literally: ONE stream of T frames (frame-serial latency) ----
    single = None
    try:
        t1 = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), 1, max_detections=D, max_tracks=MT,
                                     arm_joints=sh.arm_joints, device=local_rank)
        d1, c1 = d_dets[:1].contiguous(), d_counts[:1].contiguous()
        o1 = t1.alloc_outputs(T, nviews=False, assoc=False)
        best = time_kernel(t1, d1, c1, o1)
        t1.check()
        hd1, hc1 = h_dets[:1].numpy(), h_counts[:1].numpy()
        ho1 = dict(count=np.empty((1, T), np.int32), ids=np.empty((1, T, MT), np.int32),
                   joints=np.empty((1, T, MT, J, 3), np.float32), nviews=None, assoc=None)   # t1: default rows = max_tracks
        t1.run_host(hd1, hc1, fresh=True, nviews=False, out=ho1)
        w0 = time.perf_counter()
        for _ in range(3):
            t1.run_host(hd1, hc1, fresh=True, nviews=False, out=ho1)
        wall = (time.perf_counter() - w0) / 3
        single = {"workload": f"one {a.shape} stream of {T} frames on one thread group", "value": T / (best * 1e-3),
                  "unit": "frames/s", "us_per_frame": best * 1e3 / T, "e2e_value": T / wall, "launch": t1.launch_info()}
        t1.close()
    except Exception as e:       # the headline numbers above must not depend on this extra
        single = {"error": str(e)[:200]}

    # ---- roofline of the tracker kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_out = reports / float(S * T)
    b_frame = frame_bytes(V, D, J, n_out)
    achieved = b_frame * S * T / (kernel_ms * 1e-3) / 1e9
    traffic, secondary = None, None
    try:   # measured by ncu --set full on this launch shape (profiles/r02_traffic.json), scaled to one launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if a.shape == "shelf" and tj.get("threads_per_cta") == launch_info["threads_per_cta"] and \
                tj.get("registers_per_thread") == launch_info["registers_per_thread"]:
            traffic = tj["dram_bytes_per_frame"] * S * T
            secondary = {"pipe": "fp64", "frac": tj["fp64_pipe_pct"] / 100.0, "issue_slots_busy": tj["issue_active_pct"] / 100.0,
                         "source": tj["source"]}
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": b_frame * S * T, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "kernel": "k_track_sequences", "kernel_ms": kernel_ms, "bytes_per_frame": b_frame, "secondary": secondary,
                "note": "frame-serial FP64 state machine: latency/FP64-issue bound, not HBM bound"}

    del d_dets, d_counts, out, d_gt, h_dets, h_counts
    torch.cuda.empty_cache()

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        rate, el = cpu_sample(a.shape, a.cpu_frames, 1)
        cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"first {a.cpu_frames} frames of one {a.shape} sequence after 10 warm-up frames "
                         f"({el:.1f} s), numpy oracle = bit-identical restatement of the reference path"}

    # ---- the other BASELINE.json configurations and the per-frame drop-in API (rank 0 of a 1-GPU run) --------
    others, perframe = None, None
    if world == 1 and not a.no_extras and not strong:
        others = {}
        for name, fn in (("campus", lambda: other_config("campus", 2368, 148, dev, peak, 1500)),
                         ("panoptic", lambda: other_config("panoptic", 592, 37, dev, peak, 400)),
                         ("dense", lambda: dense_config(dev, peak))):
            try:
                others[name] = fn()
            except Exception as e:
                others[name] = {"error": repr(e)[:300]}
        try:
            perframe = per_frame_api(dev)
        except Exception as e:
            perframe = {"error": repr(e)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.shape, sh, T, S) if not strong else
                   f"{a.shape}: {sh.V} cameras x {sh.P} people x {sh.J} joints x {T} frames, {a.sequences_total} independent "
                   f"sequences in total sharded over {world} GPU(s), NCCL all-gather of the results",
                   "sequences_per_gpu": S, "frames": T, "max_tracks": MT, "max_report": MR,
                   "l2": f"inputs larger than L2 ({S * T * V * D * J * 12 / 1e9:.2f} GB of detections per GPU per step)",
                   "gen_seconds": round(gen_s, 1), "cpus_bound_per_rank": numa_cpus, "reports_per_step": int(counters[0].item()),
                   "pcp_percent": (round(100.0 * counters[2].item() / max(1, counters[3].item()), 3) if do_eval else None),
                   "mpjpe_mm": (round(counters[4].item() / max(1, counters[5].item()) / 1e3, 3) if do_eval else None),
                   "launch": launch_info, "gather_ms_per_step": gather_ms,
                   "tracker_ms_per_step": kernel_ms, "counters_ms_per_step": eval_ms},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity_sample": parity, "single_stream": single, "config5_strong_1024": config5, "other_configs": others,
        "per_frame_api": perframe,
    }
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
