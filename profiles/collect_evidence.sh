#!/bin/bash
# Runs on the GPU box (through gpurun): the round's evidence in one call.  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/ev_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ev_smoke.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/ev_bench_reference.json 2>> gpurun_out/ev_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ev_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ev_bench_under_ncu.json 2>> gpurun_out/ev_bench.err
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 12 -c 1 -o gpurun_out/ev_ncu_mlp \
    python profiles/step_probe.py 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_qan_xattn_ln -s 8 -c 1 -o gpurun_out/ev_ncu_qan \
    python profiles/step_probe.py 2 > /dev/null 2>&1
python profiles/mlp_trace.py > gpurun_out/ev_mlp_trace.txt 2>&1
python profiles/attn_trace.py > gpurun_out/ev_attn_trace.txt 2>&1
python profiles/pdl_probe.py > gpurun_out/ev_pdl.txt 2>&1
cat gpurun_out/ev_pytest.txt gpurun_out/ev_smoke.txt; cut -c1-600 gpurun_out/ev_bench.json; tail -3 gpurun_out/ev_bench.err
