#!/usr/bin/env python
"""bench.py -- synthetic multi-view frames/s (triangulated + associated) of the part-aware tracker.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], batched as configs[4]): Shelf-shaped synthetic streams -- 5
cameras, 4 people, 14 joints, 3200 frames per sequence -- S independent sequences per GPU (weak
scaling: every rank tracks its own S sequences; sequences never span GPUs).  One "step" = one pass
of the tracker over all S x T frames of the rank, starting from empty trackers.

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
    ap.add_argument("--sequences", type=int, default=1184, help="independent sequences per GPU (148 SMs x 8 CTAs; 1776 = 12 per SM is ~4% faster but needs 35 GB of host memory per rank)")
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
# GPU arm
# ----------------------------------------------------------------------------------------------
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
    S, T = a.sequences, (a.frames or sh.T)
    V, J, D = sh.V, sh.J, sh.P
    MT = 8 if sh.P <= 6 else 12
    # host-memory guard: every rank pins its detections and result buffers; never take more than half of
    # this rank's share of the free host memory (an 8-rank run must not drive the box out of memory)
    try:
        import psutil
        per_seq = T * (V * D * J * 3 * 4 + V * 4) + T * (MT * (J * 3 * 4 + 4) + 4) + T * sh.P * J * 3 * 8
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        budget = 0.5 * psutil.virtual_memory().available / max(1, local_world)
        if S * per_seq > budget:
            S = max(148, int(budget / per_seq) // 148 * 148)
    except Exception:
        pass

    # ---- synthetic input, generated straight into pinned host memory ---------------------------
    h_dets = torch.empty((S, T, V, D, J, 3), dtype=torch.float32, pin_memory=True)
    h_counts = torch.empty((S, T, V), dtype=torch.int32, pin_memory=True)
    seq_ids = [rank * S + s for s in range(S)]          # distinct seeds on every rank
    do_eval = (J == 14)                                 # PCP counters need the 14 Shelf/Campus joints
    h_gt = np.empty((S, T, sh.P, J, 3), np.float64) if do_eval else None
    t0 = time.time()
    rig = generate(a.shape, seq_ids, T, h_dets.numpy(), h_counts.numpy(), h_gt)
    gen_s = time.time() - t0
    cams = camera.GetCameraParameters(rig)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), S, max_detections=D, max_tracks=MT,
                                  arm_joints=sh.arm_joints, device=local_rank)
    d_dets = h_dets.to(dev, non_blocking=True)
    d_counts = h_counts.to(dev, non_blocking=True)
    out = trk.alloc_outputs(T, nviews=False, assoc=False)
    # run counters, summed over ranks -- the only collective of the path: [reports, frames,
    # PCP correct, PCP evaluated, MPJPE sum (um), joints counted]
    counters = torch.zeros(6, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    from pam_b200 import evaluate
    d_gt = torch.from_numpy(h_gt).to(dev) if do_eval else None
    h_gt = None                                          # the ground truth lives on the device only
    pcp = torch.zeros((sh.P, 10, 2), dtype=torch.int64, device=dev)
    mpj = torch.zeros(2, dtype=torch.float64, device=dev)

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
        pdist.reduce_counters(counters)

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize(dev)
    trk.check()
    launches0 = trk.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
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
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    reports = int(out["count"].sum().item())
    total_frames = world * S * T
    value = total_frames * a.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    e2e = None
    if not a.no_e2e:
        ho = dict(count=torch.empty((S, T), dtype=torch.int32, pin_memory=True).numpy(),
                  ids=torch.empty((S, T, MT), dtype=torch.int32, pin_memory=True).numpy(),
                  joints=torch.empty((S, T, MT, J, 3), dtype=torch.float32, pin_memory=True).numpy(),
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

    if rank != 0:
        return

    # ---- BASELINE.json configs[1] taken literally: ONE stream of T frames (frame-serial latency) ----
    single = None
    try:
        t1 = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), 1, max_detections=D, max_tracks=MT,
                                     arm_joints=sh.arm_joints, device=local_rank)
        d1, c1 = d_dets[:1].contiguous(), d_counts[:1].contiguous()
        o1 = t1.alloc_outputs(T, nviews=False, assoc=False)
        best = 1e30
        for it in range(4):
            t1.restart()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            t1.run(d1, c1, out=o1, frame0=0)
            s1.record(stream)
            torch.cuda.synchronize(dev)
            if it:
                best = min(best, s0.elapsed_time(s1))
        t1.check()
        hd1, hc1 = h_dets[:1].numpy(), h_counts[:1].numpy()
        ho1 = dict(count=np.empty((1, T), np.int32), ids=np.empty((1, T, MT), np.int32),
                   joints=np.empty((1, T, MT, J, 3), np.float32), nviews=None, assoc=None)
        t1.run_host(hd1, hc1, fresh=True, nviews=False, out=ho1)
        w0 = time.perf_counter()
        for _ in range(3):
            t1.run_host(hd1, hc1, fresh=True, nviews=False, out=ho1)
        wall = (time.perf_counter() - w0) / 3
        single = {"workload": f"one {a.shape} stream of {T} frames on one CTA", "value": T / (best * 1e-3),
                  "unit": "frames/s", "us_per_frame": best * 1e3 / T, "e2e_value": T / wall}
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
    b_frame = 4.0 * (3 * V * D * J + 3 * n_out * J + n_out)
    achieved = b_frame * S * T / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:   # measured DRAM bytes per frame of this kernel (ncu --set full, profiles/), scaled to one launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if a.shape == "shelf":
            traffic = tj["dram_bytes_per_frame"] * S * T
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes": b_frame * S * T, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "kernel": "k_track_sequences", "kernel_ms": kernel_ms, "bytes_per_frame": b_frame,
                "note": "frame-serial FP64 state machine: latency/FP64-issue bound, not HBM bound"}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        rate, el = cpu_sample(a.shape, a.cpu_frames, 1)
        cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"first {a.cpu_frames} frames of one {a.shape} sequence after 10 warm-up frames "
                         f"({el:.1f} s), numpy oracle = bit-identical restatement of the reference path"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.shape, sh, T, S), "sequences_per_gpu": S, "frames": T, "max_tracks": MT,
                   "l2": f"inputs larger than L2 ({d_dets.numel() * 4 / 1e9:.2f} GB of detections per GPU per step)",
                   "gen_seconds": round(gen_s, 1), "cpus_bound_per_rank": numa_cpus, "reports_per_step": int(counters[0].item()),
                   "pcp_percent": (round(100.0 * counters[2].item() / max(1, counters[3].item()), 3) if do_eval else None),
                   "mpjpe_mm": (round(counters[4].item() / max(1, counters[5].item()) / 1e3, 3) if do_eval else None),
                   "threads_per_cta": int(os.environ.get("PAM_TRACK_THREADS", "0")) or "auto"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "single_stream": single,
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
