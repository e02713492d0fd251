"""Per-phase instruction mix and warp-stall breakdown of a tracker-kernel ncu capture (needs -lineinfo, --import-source on).

  python tools/ncu_phases.py <report.ncu-rep> <libpam.so the capture ran> <kernel substring, e.g. 256ELi2ELi1E> <frames in the launch>

SASS instructions are attributed to the phase of frame_step (csrc/pam_track.h, "---- phase N" markers) whose source
lines they follow in the line table of the cubin (nvdisasm -g)."""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, kvar, frames = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "part-aware_measurement_for_3d_pose_estimation_and_tracking_b200", "csrc", "pam_track.h")).read().splitlines()
fs_line = next(i + 1 for i, l in enumerate(src) if "PAM_HD void frame_step" in l)
names = {1: "1 age", 2: "2 affinity", 3: "3 assign", 4: "4 gather", 5: "5 filter+dlt", 6: "6 smooth", 7: "7 lifecycle", 8: "8 init"}
bounds = [(0, "setup")]
for i, l in enumerate(src):
    m = re.search(r"---- phase (\d)", l)
    if m and i + 1 > fs_line:
        bounds.append((i + 1, names[int(m.group(1))]))
    if "do_init = (V >= 2" in l:
        bounds[-1] = bounds[-1]
end_line = next(i + 1 for i, l in enumerate(src) if "PAM_HD void persist_views" in l)
bounds.append((end_line, "tail"))
bl = [b[0] for b in bounds]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
func = None; cur = None; addr = {}
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and func and kvar in func and "k_track_sequences" in func: addr[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[1]
ia, isamp, iex, ith = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict((b[1], [0, 0, 0, 0]) for b in bounds)
sub = collections.defaultdict(int); ops = collections.Counter(); st = collections.defaultdict(collections.Counter)
base = None; phase = "setup"
for r in rows[2:]:
    if len(r) <= max(stall_cols): continue
    a = int(r[ia], 16)
    if base is None: base = a
    key = addr.get(a - base)
    if key is None: continue
    (f, n), txt = key
    if f == "pam_track.h" and n >= fs_line: phase = bounds[bisect.bisect_right(bl, n) - 1][1]
    elif f == "pam_lib.cu": phase = "setup"
    ex = int(r[iex] or 0)
    e = agg[phase]; e[0] += ex; e[1] += int(r[isamp] or 0); e[2] += 1; e[3] += int(r[ith] or 0)
    op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0]
    sub[(phase, op)] += ex; ops[op] += ex
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: st[phase][hdr[i][6:]] += v
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"{'phase':14s} {'inst/frame':>10s} {'share':>6s} {'stall%':>7s} {'static':>7s} {'thr/inst':>8s}")
for k, v in agg.items():
    print(f"{k:14s} {v[0]/frames:10.0f} {100*v[0]/tot:5.1f}% {100*v[1]/max(1,tots):6.1f}% {v[2]:7d} {v[3]/max(1,v[0]):8.1f}")
print("top opcodes overall (per frame):", ", ".join(f"{k}={v/frames:.0f}" for k, v in ops.most_common(24)))
for ph in agg:
    it = sorted(((k[1], v) for k, v in sub.items() if k[0] == ph), key=lambda x: -x[1])[:12]
    if it: print(ph, ":", ", ".join(f"{k}={v/frames:.0f}" for k, v in it))
cols = ["wait", "no_inst", "long_sb", "short_sb", "selected", "not_selected", "math", "branch_resolving", "barrier", "dispatch"]
tt = sum(sum(c.values()) for c in st.values())
print(f"\n{'phase':14s}" + "".join(f"{n[:9]:>10s}" for n in cols) + "     total")
for b in agg:
    c = st[b]
    print(f"{b:14s}" + "".join(f"{100*c[n]/max(1,tt):9.1f}%" for n in cols) + f" {100*sum(c.values())/max(1,tt):8.1f}%")
