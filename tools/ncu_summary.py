"""Turn the ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1c.csv profiles/r1c_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_r1c.ncu-rep profiles/r1c_full.md

`launches`: per-kernel launch count, total / mean device time and SHARE of the captured launches
(cold-cache, serialised: the share is what is comparable with bench.py's CUDA-event kernel_ms).
`full`: one row per captured launch with the metrics the roofline and the optimisation notes use.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
]


def short(name):
    name = name.replace("lmcma::", "").replace("void ", "")
    return name.split("(")[0]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        if r[iu] == "us":
            v *= 1e3
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list summary (%s)\n\n" % src)
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and "
                 "serialised; compare SHARES with bench.py's CUDA-event `kernel_ms`, not absolutes.\n\n")
        fh.write("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.1f | %.2f | %.1f %% |\n" % (k, n, t / 1e3, t / n / 1e3, 100 * t / total))
        fh.write("\ncaptured launches: %d, total %.1f us\n" % (sum(a[0] for a in agg.values()), total / 1e3))
        # the kernels of the generation only (k_probe is the one-off co-scheduling probe, which times out under the profiler
        # by design; the fill kernel is bench.py's L2 flush; k_brick / k_pack_pairs are set-up)
        gen = OrderedDict((k, v) for k, v in agg.items() if k.startswith(("k_cost", "k_rank", "k_update", "k_sample", "k_gate", "k_gram", "k_coef", "k_combine")))
        gtot = sum(a[1] for a in gen.values())
        if gtot > 0:
            fh.write("\nShare among the kernels of the generation (what bench.py's `kernel_ms` shares are compared with):\n\n")
            fh.write("| kernel | mean us | share of the generation |\n|---|---|---|\n")
            for k, (n, t) in sorted(gen.items(), key=lambda kv: -kv[1][1]):
                fh.write("| `%s` | %.2f | %.1f %% |\n" % (k, t / n / 1e3, 100 * t / gtot))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as fh:
        fh.write("# ncu --set full summary (%s)\n\n" % src)
        fh.write("One row per captured launch (`--clock-control none --import-source on`).  `dram_read + dram_write` is the "
                 "`traffic` bench.py reports for the dominant kernel.\n\n")
        cols = [(m, s) for m, s in FULL_METRICS if m in hdr]
        fh.write("| kernel | " + " | ".join(s for _, s in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            cells = []
            for m, _ in cols:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                cells.append("%s %s" % (v, units[i]) if units[i] and units[i] != "%" else v)
            fh.write("| `%s` | %s |\n" % (short(r[hdr.index("Kernel Name")]), " | ".join(cells)))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
