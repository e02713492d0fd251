"""Aggregate an ncu SASS source-page CSV by CUDA source line using nvdisasm -g line info.

usage: ncu_lines.py <report.ncu-rep> <lib.so> [kernel-substring]
Prints per source line: executed warp instructions and stall samples (development aid)."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# walk the disassembly: track current function, current line annotation, instruction offsets
func = None; cur = None; addr2line = {}
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = m.group(3)
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and func:
        addr2line[(func, int(m.group(1), 16))] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find kernel name row + header
kname = rows[0][1]
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
funcs = sorted({f for f, _ in addr2line})
# kernel function = the one whose mangled name contains the kernel's base name
kbase = re.match(r"(\w+)", kname.replace("void ", "")).group(1)
cands = [f for f in funcs if kbase in f]
kfunc = ([f for f in cands if os.environ.get("KVARIANT", "") in f] or cands)[0]
for r in rows[2:]:
    if len(r) <= iex: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    key = addr2line.get((kfunc, a - base))
    line = key[0] if key else ("?", 0)
    e = agg[line]
    e[0] += int(r[iex] or 0); e[1] += int(r[isamp] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: e[2][hdr[i]] += v
tot_e = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"kernel {kname[:60]}  total inst {tot_e}  samples {tot_s}")
srcs = {}
def src(f, n):
    for d in ("part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc",):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][n - 1].strip()[:90] if 0 < n <= len(srcs[p]) else ""
    return ""
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(os.environ.get("TOP", "45"))]
for (f, n), (e, s, st) in top:
    print(f"{s/tot_s*100:5.1f}% smp {e/tot_e*100:5.1f}% inst  {f}:{n:<4d} {','.join(f'{k[6:]}={v}' for k, v in st.most_common(3)):40s} | {src(f, n)}")
