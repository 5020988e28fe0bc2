"""Per-CUDA-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_update [top]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; out = []; seen_kernel = 0
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; seen_kernel += 1
        if seen_kernel > 1: break
        continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        out.append((int(d["# Samples"]), int(d["Instructions Executed"]), int(r[0]), r[1].strip()[:110], d))
    except Exception:
        pass
tot_s = sum(o[0] for o in out) or 1; tot_i = sum(o[1] for o in out) or 1
print("total samples %d, warp insts %d" % (tot_s, tot_i))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for s, i, ln, src, d in sorted(out, key=lambda o: -o[0])[:top]:
    st = sorted(((int(d[k] or 0), k[6:]) for k in stalls), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% inst  L%-4d %-110s %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, ln, src, " ".join("%s:%d" % (k, v) for v, k in st if v)))
