#!/bin/bash
mkdir -p gpurun_out
for tag in new old; do
  if [ $tag = old ]; then export IPC_B200_LIB=/root/repo/scratch_ab/libold.so; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python scripts/ab2.py > gpurun_out/ncu_$tag.log 2>&1
  python - <<PY
import csv
lines=[l for l in open('gpurun_out/launches_$tag.csv') if not l.startswith('==')]
n=0
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    if 'chain' in row['Kernel Name'] and n<6: print("$tag %10.3f ms %s %s %s"%(v/1e6,row['Kernel Name'][:50],row['Grid Size'],row['Block Size'])); n+=1
PY
done
