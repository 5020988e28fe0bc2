python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/s5_bench_n2.log 2> gpurun_out/s5_bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/s5_bench_n2.log') if x.startswith('{')][-1]
j=json.loads(l); print('n_gpus', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/c4_split_population.py 256 500 8192 30 > gpurun_out/s5_c4_n2.log 2>&1; echo "c4 rc=$?"; tail -4 gpurun_out/s5_c4_n2.log
