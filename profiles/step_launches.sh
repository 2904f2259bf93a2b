#!/bin/bash
# launch list of the sampling step (ncu, cold caches) + the bench line
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/$1_bench_under_ncu.json 2> gpurun_out/$1_ncu.err
python profiles/summarize_launches.py gpurun_out/$1_launches.csv | tail -14
