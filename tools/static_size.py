"""Static SASS instruction count per source line / per source function region of one kernel."""
import os, re, subprocess, sys, tempfile, collections
so, kern = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
func = None; cur = None; cnt = collections.Counter(); tot = 0
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if func and kern in func and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        cnt[cur] += 1; tot += 1
print("total SASS", tot, "=", tot * 16 // 1024, "KB")
srcs = {}
def src(f, n):
    p = os.path.join("part-aware_measurement_for_3d_pose_estimation_and_tracking_b200/csrc", f)
    if os.path.exists(p):
        if p not in srcs: srcs[p] = open(p).read().splitlines()
        return srcs[p][n - 1].strip()[:90] if 0 < n <= len(srcs[p]) else ""
    return ""
for (f, n), c in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(f"{c:5d}  {f}:{n:<4d} | {src(f, n)}")
