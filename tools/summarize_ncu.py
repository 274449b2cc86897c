"""Turn gpurun_out/launches.csv and .ncu-rep captures into small tracked summaries under profiles/.
usage: summarize_ncu.py <tag> [report.ncu-rep:name ...]"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

def short(n):
    n = re.sub(r"\((CUtensorMap_st|const |float|__nv_bfloat16|int|long).*", "", n)
    return n.replace("<unnamed>::", "").replace("void ", "")[:80]

lc = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lc) and len(sys.argv) <= 2:
    lines = [l for l in open(lc) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for d in csv.DictReader(lines):
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1.0)
        k = short(d["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list, a window of {sum(cnt.values())} consecutive launches (about 1.5 train steps, B=8/GPU) inside the timed region: {T/1e6:.2f} ms serialised\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 2 --warmup 3 "
                "--no-cpu-baseline`; cold-cache, serialised - compare shares, not absolutes.\n\n| ms | share | launches | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1])[:45]:
            f.write(f"| {v/1e6:.3f} | {100*v/T:.1f}% | {cnt[k]} | `{k}` |\n")
    print("wrote launches summary")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
for spec in sys.argv[2:]:
    rep, name = spec.split(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(w) for w in want if w in hdr]
    seen = collections.Counter()
    with open(os.path.join(out_dir, f"{tag}_ncu_{name}.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["kernel"] + [hdr[i] for i in idx])
        wr.writerow(["unit"] + [rows[1][i] for i in idx])
        for r in rows[2:]:
            k = short(r[hdr.index("Kernel Name")])
            key = (k, r[hdr.index("launch__grid_size")], r[hdr.index("dram__bytes_read.sum")][:4])
            seen[key] += 1
            if seen[key] > 2:          # repeated identical launches: keep two
                continue
            wr.writerow([k] + [r[i] for i in idx])
    print("wrote", name, len(rows) - 2, "launches")
