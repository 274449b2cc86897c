"""dev tool: top stall sites of one kernel from `ncu --page source --csv` (SASS view)."""
import csv, collections, subprocess, sys
rep, kern = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the dump is a sequence of blocks: ["Kernel Name", name], header row, data rows
blocks, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks:
    if kern and kern not in b["name"]:
        continue
    hdr, data = b["hdr"], b["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    n = lambda r, k: int(float(r[ix[k]] or 0))
    tot = sum(n(r, "# Samples") for r in data) or 1
    print("==", b["name"][:90], "samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -n(r, "# Samples"))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        st = sorted([(n(r, h), h[6:]) for h in stalls], reverse=True)[:2]
        print(f"{n(r,'# Samples'):6d} {100*n(r,'# Samples')/tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].strip().split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        agg[op.split(".")[0]] += n(r, "# Samples"); cnt[op.split(".")[0]] += n(r, "Instructions Executed")
    print("-- by opcode")
    for k, v in agg.most_common(14):
        print(f"{k:12s} samples {v:6d} {100*v/tot:5.1f}%  executed {cnt[k]}")
    break
