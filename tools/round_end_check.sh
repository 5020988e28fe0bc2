#!/bin/bash
# What a round ends with on one B200 (run through gpurun from the repository root; ~6 GPU-minutes):
#   gpurun --timeout 1500 -- 'bash tools/round_end_check.sh r2'
# tests -> smoke -> both bench arms -> ncu launch list of the bench command and --set full captures of the C2 step.  Outputs
# land in gpurun_out/<tag>_*; summarise them into profiles/ with tools/ncu_summary.py.  Under ncu the co-scheduling probe fails
# (kernels are serialised), so the library builds the linear generation graph by itself.
tag=${1:-round}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 20 --warmup 3 --skip c3,c4,cpu > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on --launch-skip 184 -c 4 -f -o gpurun_out/prof_${tag} \
    python tools/profile_step.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/${tag}_ncu_full.log
