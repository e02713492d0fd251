"""Static SASS instruction count per source line for the hot part of a kernel (development aid).
usage: code_size.py <report.ncu-rep> <lib.so> <frames>"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so, frames = sys.argv[1], sys.argv[2], float(sys.argv[3])
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
func = None; cur = None; addr2line = {}
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and func: addr2line[(func, int(m.group(1), 16))] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]; hdr = rows[1]
ia, iex = hdr.index("Address"), hdr.index("Instructions Executed")
kbase = re.match(r"(\w+)", kname.replace("void ", "")).group(1)
cands = [f for f in sorted({f for f, _ in addr2line}) if kbase in f]
kfunc = ([f for f in cands if os.environ.get("KVARIANT", "ILi128ELi4") in f] or cands)[0]
base = None
hot = collections.Counter(); ops = collections.Counter()
for r in rows[2:]:
    if len(r) <= iex: continue
    a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None: base = a
    e = int(r[iex] or 0)
    if e < frames * 0.5: continue
    key = addr2line.get((kfunc, a - base))
    if key:
        hot[key[0]] += 1
        ops[key[1].split()[0].split(".")[0]] += 1
print("hot SASS instructions:", sum(hot.values()))
srcs = {}
def src(f, n):
    p = os.path.join("part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc", f)
    if os.path.exists(p):
        if p not in srcs: srcs[p] = open(p).read().splitlines()
        return srcs[p][n - 1].strip()[:80] if 0 < n <= len(srcs[p]) else ""
    return ""
for (f, n), c in hot.most_common(40):
    print(f"{c:5d}  {f}:{n:<4d} | {src(f, n)}")
print(ops.most_common(25))
