"""Bucket the PC samples of one kernel from an .ncu-rep by CUDA source line ranges.
  python tools/ncu_phases.py rep kernel 'name:lo-hi' ..."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
buckets = []
for spec in sys.argv[3:]:
    name, rng = spec.split(":"); lo, hi = rng.split("-"); buckets.append((name, int(lo), int(hi)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; tot = {}; inst = {}; seen = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; seen += 1
        if seen > 1: break
        continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    d = dict(zip(hdr, r))
    try: ln, smp, ins = int(r[0]), int(d["# Samples"]), int(d["Instructions Executed"])
    except Exception: continue
    for name, lo, hi in buckets:
        if lo <= ln <= hi:
            tot[name] = tot.get(name, 0) + smp; inst[name] = inst.get(name, 0) + ins; break
    else:
        tot["other"] = tot.get("other", 0) + smp; inst["other"] = inst.get("other", 0) + ins
S = sum(tot.values()) or 1; I = sum(inst.values()) or 1
for k in tot: print("%-12s %5.1f%% samples  %5.1f%% warp insts" % (k, 100.0 * tot[k] / S, 100.0 * inst[k] / I))
