python -m pytest tests -m gpu -q > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s5_pytest.log
python tools/c3_batched.py 512 50 > gpurun_out/s5_c3.log 2>&1; tail -3 gpurun_out/s5_c3.log
python bench.py --steps 300 --warmup 5 > gpurun_out/s5_bench.log 2> gpurun_out/s5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/s5_bench.log') if x.startswith('{')][-1]
j=json.loads(l); print(j['value'], j['ms_per_step'], j['steady_state_l2_warm']['ms_per_step'], j['kernel_ms'], j['e2e']['ms_per_step'], j['e2e']['calls_ms_rank0'])
PY
