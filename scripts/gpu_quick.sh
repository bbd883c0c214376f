#!/bin/bash
# quick GPU pass: tests, launch list on a 100k sample, short bench on a 400k sample
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --checks 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    if 'chain' in row['Kernel Name']: print("%10.3f ms %s %s %s"%(v/1e6,row['Kernel Name'][:50],row['Grid Size'],row['Block Size']))
PY
timeout 900 python bench.py --checks 400000 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_400k.json 2> gpurun_out/bench_400k.err; cut -c1-400 gpurun_out/bench_400k.json; tail -3 gpurun_out/bench_400k.err
K=${1:-"chain_check_se2<256"}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_check_se2" -s 3 -c 2 -f -o gpurun_out/prof python bench.py --checks 20000 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/prof.ncu-rep
