"""Launch-shape sweep of the tracker kernel on one GPU (development aid; results go to profiles/).

  python tools/sweep_shapes.py [--unique 296] [--frames 3200] [--out gpurun_out/sweep.json] shape@S ...

Every argument `G:Q:regs[:lean]@S` (or `auto@S`) runs S Shelf-shaped sequences (a tiling of `--unique` seeded ones)
through pam_track_sequences with that PAM_TRACK_SHAPE and reports kernel ms (CUDA events, best of 3 after a
warm-up) and frames/s; outputs of the unique sequences are compared with the first configuration's
(bit-identical or the run is flagged)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pam_b200  # noqa
from pam_b200 import camera, synth, tracker

ap = argparse.ArgumentParser()
ap.add_argument("--unique", type=int, default=296)
ap.add_argument("--frames", type=int, default=3200)
ap.add_argument("--shape", default="shelf")
ap.add_argument("--max-tracks", type=int, default=8)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
ap.add_argument("configs", nargs="+")
a = ap.parse_args()

sh = synth.SHAPES[a.shape]
from concurrent.futures import ThreadPoolExecutor
rig = synth.make_rig(a.shape)
t0 = time.time()
with ThreadPoolExecutor(min(32, os.cpu_count() or 8)) as ex:
    streams = list(ex.map(lambda s: synth.make_stream(a.shape, s, a.frames, rig=rig), range(a.unique)))
print(f"generated {a.unique} sequences in {time.time() - t0:.1f} s", flush=True)
dets_u = torch.from_numpy(np.stack([s.dets for s in streams])).cuda()
counts_u = torch.from_numpy(np.stack([s.counts for s in streams])).cuda()
cams = camera.GetCameraParameters(rig)
D = dets_u.shape[3]
ref = None
rows = []
try:
    import pynvml
    pynvml.nvmlInit()
    _nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    def clocks():
        return dict(sm_mhz=pynvml.nvmlDeviceGetClockInfo(_nv, pynvml.NVML_CLOCK_SM), power_w=pynvml.nvmlDeviceGetPowerUsage(_nv) / 1000.0)
except Exception:
    def clocks():
        return {}
cache = {}
for cfg in a.configs:
    shp, S = cfg.split("@")
    S = int(S)
    # "<shape>/c": the sequences of a CTA start every frame together; "/p": ... every phase together
    os.environ["PAM_TRACK_CONVOY"] = "1" if shp.endswith("/c") else ("2" if shp.endswith("/p") else "0")
    shp = shp[:-2] if shp[-2:] in ("/c", "/p") else shp
    if shp == "auto":
        os.environ.pop("PAM_TRACK_SHAPE", None)
    else:
        os.environ["PAM_TRACK_SHAPE"] = shp
    if S not in cache:
        cache.clear()
        torch.cuda.empty_cache()
        reps = (S + a.unique - 1) // a.unique
        cache[S] = (dets_u.repeat(reps, 1, 1, 1, 1, 1)[:S].contiguous(), counts_u.repeat(reps, 1, 1)[:S].contiguous())
    dets, counts = cache[S]
    trk = tracker.SequenceTracker(cams, synth.tracker_params(a.shape), S, max_detections=D, max_tracks=a.max_tracks,
                                  arm_joints=sh.arm_joints)
    try:
        info = trk.launch_info()
    except Exception as e:
        print(f"{cfg}: {e}", flush=True)
        rows.append(dict(config=cfg, error=str(e)))
        continue
    out = trk.alloc_outputs(a.frames, nviews=False, assoc=False)
    best = 1e30
    clk = {}
    for it in range(4):
        trk.restart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trk.run(dets, counts, out=out, frame0=0)
        e1.record()
        if it == 3:
            time.sleep(0.03)
            clk = clocks()          # sampled while the kernel runs
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    st = trk.check(strict=False)
    nu = min(a.unique, S)
    got = (out["count"][:nu].cpu().numpy(), out["ids"][:nu].cpu().numpy(), out["joints"][:nu].cpu().numpy())
    k = got[0]
    mask = np.arange(got[1].shape[2])[None, None, :] < k[:, :, None]
    sig = (k.copy(), np.where(mask, got[1], 0), np.where(mask[..., None, None], got[2], 0))
    same = None
    if ref is None:
        ref = sig
    else:
        m = min(len(ref[0]), nu)
        same = all(np.array_equal(x[:m], y[:m]) for x, y in zip(ref, sig))
    fps = S * a.frames / (best * 1e-3)
    row = dict(config=cfg, S=S, ms=best, mfps=fps / 1e6, info=info, identical_to_first=same, warn=int((st != 0).sum()), clocks=clk,
               reports=int(k.sum()))
    rows.append(row)
    print(json.dumps(row), flush=True)
    trk.close()
    del out, trk
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(rows, open(a.out, "w"), indent=1)
