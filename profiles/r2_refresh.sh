#!/bin/bash
# Refresh of the numbers that depend on the step kernels (after the last kernel change of the round); r2_collect.sh = the full set.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/ev2_bench_c2.json 2> gpurun_out/ev2_bench.err
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/ev2_bench_c3.json 2>> gpurun_out/ev2_bench.err
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/ev2_bench_c4.json 2>> gpurun_out/ev2_bench.err
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu > gpurun_out/ev2_bench_c5.json 2>> gpurun_out/ev2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/ev2_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>> gpurun_out/ev2_bench.err
ncu --set full --clock-control none --import-source on -k regex:^k_attn_xattn_ln -s 1 -c 1 -o gpurun_out/ev2_ncu_k_attn_xattn_ln python profiles/step_probe.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 12 -c 1 -o gpurun_out/ev2_ncu_mlp python profiles/step_probe.py 2 > /dev/null 2>&1
python profiles/mlp_trace.py > gpurun_out/ev2_mlp_trace.txt 2>&1
python profiles/chain_probe.py 64 12 loop > gpurun_out/ev2_chain_probe.txt 2>&1
python profiles/chain_probe.py 8 12 loop > gpurun_out/ev2_chain_probe_b8.txt 2>&1
python profiles/gemm_trace.py > gpurun_out/ev2_gemm_trace.txt 2>&1
python profiles/attn_trace.py > gpurun_out/ev2_attn_timeline.txt 2>&1
python profiles/r2_probe.py --config3 > gpurun_out/ev2_probe.txt 2>&1
cut -c1-200 gpurun_out/ev2_bench_c2.json; tail -2 gpurun_out/ev2_bench.err
