"""profiles/roofline_r1.json from an ncu CSV of (gpu__time_duration, dram bytes) per lfm_dgemm_kernel launch."""
import csv, json, re, sys
src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
per = {}
for r in csv.DictReader(lines):
    k = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]))
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if "byte" in u:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per.setdefault(k, {})[r["Metric Name"]] = v
# one evaluation = the launches after the last lfm_residual_kernel (the first kernel of an NLML evaluation), if any
starts = sorted(int(i) for (i, n) in per if "lfm_residual_kernel" in n)
first = starts[-1] if starts else -1
per = {k: v for k, v in per.items() if int(k[0]) >= first and "lfm_dgemm_kernel" in k[1]}
bulk = [v for (i, n), v in per.items() if "<0, 1, 1, 4, 2" not in n]
chain = [v for (i, n), v in per.items() if "<0, 1, 1, 4, 2" in n]
tot = sum(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in bulk)
secs = sum(v.get("gpu__time_duration.sum", 0.0) for v in bulk) * 1e-9
out = {"source": src, "launches_bulk": len(bulk), "launches_chain": len(chain), "dgemm_seconds_serialised": secs,
       "dgemm_dram_bytes_per_launch": tot / max(len(bulk), 1), "dgemm_dram_bytes_per_eval": tot,
       "note": "dram__bytes_read.sum + dram__bytes_write.sum over the bulk lfm_dgemm_kernel launches of ONE N=4000 NLML+grad "
               "evaluation (ncu serialises launches; L2 holds most of the 2 x 134 MB working set, so this is far below the "
               "bytes the tiles request)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
