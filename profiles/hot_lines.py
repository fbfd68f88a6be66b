#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm line info to attribute executed warp instructions
to CUDA source lines.  usage: hot_lines.py <report.ncu-rep> <kernel regex> <mangled-name substring> [top]"""
import csv, re, subprocess, sys, collections, os, tempfile

rep, kre, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
skip = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # kernel instances (in report order) to skip
name_has = sys.argv[6] if len(sys.argv) > 6 else ""    # e.g. "(bool)1": only instances whose demangled name contains this
# (the source page lists every instance twice; template instances share one base name, so a wrong `skip` silently
#  pairs one instance's counters with another's line table -- select by name)
so = os.environ.get("NUTDB_SO") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nutdb_b200", "libnutdb_gpu.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("nutdb_gpu.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
cur_fn, cur_line = None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1); cur_line = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur_fn and mangled in cur_fn:
        line_of[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; base = None; acc = collections.Counter(); samples = collections.Counter(); total = 0; started = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if name_has and name_has not in r[1]:
            if started: break
            hdr = None; continue
        if skip > 0:
            skip -= 1; hdr = None; continue
        if started: break
        started = True; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if started and hdr and len(r) == len(hdr):
        a = int(r[0], 16)
        if base is None: base = a
        ie = float(r[hdr.index("Instructions Executed")]); sm = float(r[hdr.index("# Samples")])
        key = line_of.get(a - base)
        acc[key] += ie; samples[key] += sm; total += ie
print("total warp instructions", total)
src_cache = {}
def src(key):
    if not key: return ""
    f, l = key
    for d in ("nutdb_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:100] if l - 1 < len(src_cache[p]) else ""
    return ""
stot = sum(samples.values())
for key, v in acc.most_common(top):
    print("%5.1f%% inst %5.1f%% samples  %s:%s  %s" % (100 * v / total, 100 * samples[key] / max(stot, 1), key[0] if key else None, key[1] if key else None, src(key)))
