"""Aggregate the ncu source page (per-SASS-instruction warp-stall samples) of one kernel per PHASE of its CUDA source.

  python tools/ncu_phases.py gpurun_out/prof_team_s4.ncu-rep dis_project_b200/csrc/build/batched_warp.o \
         _Z23lfm_batched_warp_kernelILi4EEv11BatchedArgsi dis_project_b200/csrc/batched_warp.cu profiles/out.md "<note>"

`ncu --page source --csv` lists SASS addresses; `nvdisasm -g` of the same cubin gives address -> (file, line); an
instruction inlined from another file is attributed to the last line of the kernel's own file before it.  Phases are the
`// ---- ` comment lines of the source."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, mangled, srcfile, dst = sys.argv[1:6]
note = sys.argv[6] if len(sys.argv) > 6 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + mangled + ":")][0]
amap, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        if amap:
            break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m:
        amap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
H = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
src = open(srcfile).read().split("\n")
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ---- ")]
def phase(ln):
    p = "(before the first marker)"
    for m, t in marks:
        if ln >= m:
            p = t
    return p
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict()
tot = collections.Counter()
last = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    loc = amap.get(int(r[0], 16) - base)
    if loc and os.path.basename(loc[0]) == os.path.basename(srcfile):
        last = loc[1]
    c = agg.setdefault(phase(last), collections.Counter())
    smp, ins = int(r[H["# Samples"]] or 0), int(r[H["Instructions Executed"]] or 0)
    c["samples"] += smp; c["inst"] += ins; tot["samples"] += smp; tot["inst"] += ins
    for s in stall:
        v = int(r[H[s]] or 0); c[s] += v; tot[s] += v
with open(dst, "w") as f:
    f.write(f"# ncu source page aggregated per phase ({note})\n\nSource: `{rep}` (`ncu --set full --import-source on`), kernel `{mangled}`; "
            f"{tot['samples']} warp-stall samples, {tot['inst']} warp instructions executed.\n\nAll samples by reason: "
            + ", ".join(f"{s.replace('stall_', '')} {100 * tot[s] / tot['samples']:.1f} %" for s in sorted(stall, key=lambda s: -tot[s]) if tot[s] > 0.01 * tot["samples"])
            + "\n\n| samples | instructions | phase | top stall reasons (share of the phase's samples) |\n|---:|---:|---|---|\n")
    for ph, c in agg.items():
        top = sorted(((c[s], s.replace("stall_", "")) for s in stall), reverse=True)[:4]
        f.write(f"| {100 * c['samples'] / tot['samples']:.1f} % | {100 * c['inst'] / tot['inst']:.1f} % | `{ph[:90]}` | "
                + ", ".join(f"{k} {100 * v / max(c['samples'], 1):.0f} %" for v, k in top) + " |\n")
print(open(dst).read())
