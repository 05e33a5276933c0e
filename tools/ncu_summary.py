"""Summarise an .ncu-rep of the physics kernel: headline counters, stall mix, opcode mix, per-function share.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [build/nmp_kernels_fast.o] [kernel-substring]"""
import collections, csv, io, os, re, subprocess, sys, tempfile, bisect
rep = sys.argv[1]
obj = sys.argv[2] if len(sys.argv) > 2 else "noahmp_b200/csrc/build/nmp_kernels_fast.o"
kern = sys.argv[3] if len(sys.argv) > 3 else "land_kernelIN3nmp6OptSetILi2"
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ["Duration", "Elapsed Cycles", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Achieved Occupancy",
        "Avg. Active Threads Per Warp", "Avg. Not Predicated Off", "L1/TEX Hit Rate", "L2 Hit Rate", "DRAM Throughput",
        "Memory Throughput", "Warp Cycles Per Issued", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler",
        "Executed Instructions  ", "Stack Size", "Compute (SM) Throughput"]
for l in det.splitlines():
    if any(k in l for k in keys) and "OPT" not in l and "INF" not in l:
        print(l.rstrip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) >= 3:
    h = rr[0]
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "smsp__inst_executed.sum",
                 "sm__inst_executed_pipe_xu.sum", "smsp__thread_inst_executed.sum"):
        if name in h:
            print(f"{name:40s} {rr[-1][h.index(name)]} {rr[1][h.index(name)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{s[6:]} {v / tot:.3f}" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
ex = sum(int(r[ix["Instructions Executed"]]) for r in data)
nw = int(data[0][ix["Instructions Executed"]])
print(f"warp instructions {ex}  per warp {ex / nw:.0f}  static {len(data)}")
op = collections.Counter()
for r in data:
    o = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
    op[(o[0] if o else "?").split(".")[0]] += int(r[ix["Instructions Executed"]])
print("opcode mix:", ", ".join(f"{o} {v / ex:.3f}" for o, v in op.most_common(16)))
# per function
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
infun, cur, ins = False, None, []
for l in dis:
    if l.startswith(".text."):
        infun = kern in l
        continue
    if not infun:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined at" in m.group(3) and cur is not None:
            continue
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
if len(ins) != len(data):
    print("disassembly does not match the profiled binary", len(ins), len(data)); sys.exit()
funcs = {}
for f in set(x[0] for x in ins if x):
    p = os.path.join(os.path.dirname(obj), "..", f)
    if not os.path.exists(p):
        continue
    starts = []
    for i, l in enumerate(open(p), 1):
        m = re.match(r"(?:NMP_DEV|NMP_HD|__global__|__device__|inline|static)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(", l)
        if m:
            starts.append((i, m.group(1)))
    funcs[f] = starts
def fn(loc):
    if not loc: return "?"
    f, l = loc
    s = funcs.get(f)
    if not s: return f
    k = bisect.bisect_right([a for a, _ in s], l) - 1
    return s[k][1] if k >= 0 else f
exf, smf, lsb = collections.Counter(), collections.Counter(), collections.Counter()
for i, r in enumerate(data):
    k = fn(ins[i])
    exf[k] += int(r[ix["Instructions Executed"]]); smf[k] += int(r[ix["# Samples"]])
print("%-22s %9s %7s %7s" % ("function", "dyn/warp", "dyn%", "samp%"))
for k, v in exf.most_common(24):
    print("%-22s %9.0f %7.3f %7.3f" % (k, v / nw, v / ex, smf[k] / tot))
# SIMT efficiency per function
te, we = collections.Counter(), collections.Counter()
for i, r in enumerate(data):
    k = fn(ins[i])
    we[k] += int(r[ix["Instructions Executed"]]); te[k] += int(r[ix["Predicated-On Thread Instructions Executed"]])
print("SIMT efficiency (predicated-on threads / 32) of the heaviest functions:")
for k, v in we.most_common(16):
    print("  %-20s %.2f" % (k, te[k] / (32.0 * v)))
print("  overall %.3f" % (sum(te.values()) / (32.0 * sum(we.values()))))
# where the long-scoreboard and barrier stalls are
for stall in ("stall_long_sb", "stall_barrier", "stall_wait"):
    byf = collections.Counter()
    for i, r in enumerate(data):
        byf[fn(ins[i])] += int(r[ix[stall]])
    tot_s = sum(byf.values()) or 1
    print(stall, "by function:", ", ".join(f"{k} {v / tot_s:.2f}" for k, v in byf.most_common(10)))
# hottest source lines (innermost inlined location) by stall samples
byl = {s: collections.Counter() for s in ("stall_long_sb", "stall_wait", "# Samples")}
for i, r in enumerate(data):
    for s in byl:
        byl[s][ins[i]] += int(r[ix[s]])
for s, c in byl.items():
    t = sum(c.values()) or 1
    print(f"top lines by {s}:")
    for loc, v in c.most_common(14):
        if loc:
            print("   %-20s:%-5d %-14s %.3f" % (loc[0], loc[1], fn(loc), v / t))
