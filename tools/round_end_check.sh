#!/bin/bash
# What a round ends with on one B200 (run through gpurun from the repository root; ~5 GPU-minutes):
#   gpurun --timeout 1200 -- 'bash tools/round_end_check.sh r1f'
# tests -> smoke -> both bench arms -> ncu launch list and --set full captures of the C2 step.  Outputs land in
# gpurun_out/<tag>_*; summarise them into profiles/ with tools/ncu_summary.py.  ncu serialises kernels, so the
# captures run with LMCMA_B200_OVERLAP=0 (no side branch in the fused generation).
tag=${1:-round}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/${tag}_bench.json
LMCMA_B200_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu launch list rc=$?"
LMCMA_B200_OVERLAP=0 ncu --set full --clock-control none --import-source on --launch-skip 180 -c 8 -f -o gpurun_out/prof_${tag} \
    python tools/profile_step.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu_full.log
# open measurement from round 1: k_cost at six CTAs per SM on the many-wave batched shape (C3); the single
# query was measured and kept at seven (profiles/r1g_minb_compare.txt)
for minb in 7 6; do LMCMA_B200_COST_MINB=$minb python tools/c3_batched.py 1024 30 > gpurun_out/${tag}_c3_minb${minb}.txt 2>&1; echo "c3 minb=$minb rc=$?"; tail -2 gpurun_out/${tag}_c3_minb${minb}.txt; done
