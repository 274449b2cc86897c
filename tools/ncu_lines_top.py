"""dev tool: top CUDA source lines by stall samples (first kernel in the report matching argv[2])."""
import csv, subprocess, sys, os
rep, kern, topn = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, cur_fn, hdr, out, seen_fn = None, None, None, [], set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if len(r) == 2 and r[0] == "Function Name": cur_fn = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not cur_fn or kern not in cur_fn: continue
    if r[0] == "": continue
    ix = {h: i for i, h in enumerate(hdr)}
    si = hdr.index("# Samples")
    try: s = int(float(r[si]))
    except ValueError: continue
    stalls = [(int(float(r[i] or 0)), h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "-")]
    out.append((s, cur_file, r[0], r[1].strip()[:80], sorted(stalls, reverse=True)[:2]))
tot = sum(o[0] for o in out) or 1
print("samples", tot)
for s, f, ln, src, st in sorted(out, reverse=True)[:topn]:
    print(f"{s:6d} {100*s/tot:5.1f}% {f}:{ln:>4s} {src:80s} {st}")
