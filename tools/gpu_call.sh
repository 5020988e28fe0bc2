python -m pytest tests -m gpu -q -x -k "gram or teacher or split_population or long_free" > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s5_pytest.log
python tools/c4_kernels.py 256 1024 > gpurun_out/s5_c4.log 2>&1; tail -2 gpurun_out/s5_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 520 -c 35 --csv --log-file gpurun_out/launches_c4b.csv python tools/c4_kernels.py 256 1024 > gpurun_out/s5_c4_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_c4b.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-7:]: print(r[4][:40], r[7], r[8], r[-1])
PY
