"""Randomised configurations through the HOST build of the kernel source (tests/hostemu) against the oracle, in a
process pool: the CPU-side companion of tools/gpu_fuzz.py (same seeds, tests/test_fuzz_hostemu.py `_case` /
`_large_case`), usable where no GPU is at hand.   python tools/host_fuzz.py --seeds 0 1200 --large 0 150 [--twopass 1]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def job(arg):
    kind, seed, twopass = arg
    os.environ["PAM_HOSTEMU_TWOPASS"] = str(twopass)
    from tests import util
    from tests import test_fuzz_hostemu as tf
    from oracle import generic
    import pam_b200  # noqa
    from pam_b200 import _capi, synth
    shape, params, kw, min_valid = (tf._case if kind == "case" else tf._large_case)(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    cfg = _capi.make_config(params, shape.V, st.dets.shape[2], 32, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
    out = util.run_hostemu([st], cfg)
    oo, oa, _ = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    try:
        worst = util.compare_with_oracle(out, 0, st, oo, oa)
        return kind, seed, int(out["status"][0]), worst, ""
    except AssertionError as e:
        return kind, seed, int(out["status"][0]), -1.0, str(e)[:200]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, nargs=2, default=[0, 200])
    ap.add_argument("--large", type=int, nargs=2, default=[0, 0])
    ap.add_argument("--twopass", type=int, default=1)
    ap.add_argument("--log", default=None)
    a = ap.parse_args()
    from concurrent.futures import ProcessPoolExecutor
    from tests.hostemu import build as hb
    hb.build()
    jobs = [("case", s, a.twopass) for s in range(*a.seeds)] + [("large", s, a.twopass) for s in range(*a.large)]
    t0 = time.time()
    bad, warn, worst = [], 0, 0.0
    with ProcessPoolExecutor(max(1, (os.cpu_count() or 2) - 1)) as ex:
        for kind, seed, status, w, msg in ex.map(job, jobs, chunksize=4):
            if msg:
                bad.append((kind, seed, status, msg))
            warn += 1 if status else 0
            worst = max(worst, w)
    line = (f"TOTAL {len(jobs)} configurations on the host build (two-pass affinity {a.twopass}): {len(jobs) - len(bad)} identical to the "
            f"oracle ({warn} with a status/warning word), {len(bad)} mismatches {bad[:5]}; max |dX| {worst:.3e} m; {time.time() - t0:.0f} s")
    print(line)
    if a.log:
        open(a.log, "a").write(line + "\n")
