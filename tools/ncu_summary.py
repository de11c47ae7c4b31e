"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/launches_rX.md "<note>"
  python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep profiles/dgemm_rX.md "<note>"
"""
import collections, csv, io, re, subprocess, sys

def us(row):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else (v * 1e6 if u in ("s", "second") else v))

def launches(src, dst, note):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "") for r in rows]
    # one evaluation = from one residual kernel (first launch of an NLML evaluation) to the next
    idx = [i for i, n in enumerate(names) if "lfm_residual_kernel" in n]
    s, e = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(rows))
    agg = collections.OrderedDict(); tot = 0.0
    for r, n in zip(rows[s:e], names[s:e]):
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us(r); tot += us(r)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({note})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` of `bench.py`; "
                f"one NLML+grad evaluation (launches {s}..{e} of {len(rows)} captured). Per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n[:80]}` | {c} | {t:.1f} | {t / tot:.3f} | {t / c:.1f} |\n")
        f.write(f"| **total** | {e - s} | {tot:.1f} | 1.000 | |\n")
    print(open(dst).read())

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active", "gpu__dram_throughput",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem", "smsp__inst_executed.sum", "sm__pipe_fp64",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts", "smsp__warp_issue_stalled", "lts__t_sector_hit_rate", "sm__cycles_elapsed.avg ",
        "sm__inst_executed_pipe_fp64", "smsp__issue_active.avg.pct", "launch__occupancy_limit", "sm__maximum_warps"]

def full(src, dst, note):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full ({note})\n\nSource: `{src}` read with `ncu -i ... --page raw --csv`.\n\n")
        for d in data:
            name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name[:100]}  grid={d[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for i, h in enumerate(hdr):
                if any(k in h for k in KEYS) and d[i] not in ("", "n/a"):
                    f.write(f"| {h} | {d[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:6000])

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
