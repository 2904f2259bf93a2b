#!/bin/bash
# Round-2 evidence in one call (GPU box, through gpurun).  Outputs under gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/ev2_smoke.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/ev2_bench_c2.json 2> gpurun_out/ev2_bench.err
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/ev2_bench_c3.json 2>> gpurun_out/ev2_bench.err
python bench.py --config 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/ev2_bench_c4.json 2>> gpurun_out/ev2_bench.err
python bench.py --config 5 --steps 3 --warmup 3 --no-cpu > gpurun_out/ev2_bench_c5.json 2>> gpurun_out/ev2_bench.err
python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/ev2_bench_reference.json 2>> gpurun_out/ev2_bench.err
# launch lists (cold caches, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/ev2_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>> gpurun_out/ev2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ev2_launches_corr.csv python profiles/corr_step_only.py > /dev/null 2>> gpurun_out/ev2_bench.err
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:mlp_fused_kernel -s 12 -c 1 -o gpurun_out/ev2_ncu_mlp python profiles/step_probe.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lbs_skin_sparse -c 1 -o gpurun_out/ev2_ncu_skin python profiles/lbs_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_split_f16 -c 1 -o gpurun_out/ev2_ncu_lbsgemm python profiles/lbs_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_signed_nn_pruned -c 1 -o gpurun_out/ev2_ncu_nn python profiles/corr_step_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_projector -c 1 -o gpurun_out/ev2_ncu_proj python profiles/corr_step_only.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_qan_xattn_ln -s 4 -c 1 -o gpurun_out/ev2_ncu_k_qan_xattn_ln python profiles/step_probe.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_attn_xattn_ln -s 1 -c 1 -o gpurun_out/ev2_ncu_k_attn_xattn_ln python profiles/step_probe.py 2 > /dev/null 2>&1
# the folded QKV projection is the only GEMM on 192-column tiles
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gemm_split_f16_kernelILi192 -s 1 -c 1 -o gpurun_out/ev2_ncu_qkvgemm python profiles/step_probe.py 2 > /dev/null 2>&1
python profiles/mlp_trace.py > gpurun_out/ev2_mlp_trace.txt 2>&1
python profiles/chain_probe.py 64 12 loop > gpurun_out/ev2_chain_probe.txt 2>&1
python profiles/chain_probe.py 8 12 loop > gpurun_out/ev2_chain_probe_b8.txt 2>&1
interdiff_b200/build/dsmem_bench > gpurun_out/ev2_dsmem_bench.txt 2>&1
python profiles/gemm_trace.py > gpurun_out/ev2_gemm_trace.txt 2>&1
python profiles/attn_trace.py > gpurun_out/ev2_attn_timeline.txt 2>&1
python profiles/fused_layer_probe.py > gpurun_out/ev2_fused_layer_probe.txt 2>&1
python profiles/r2_probe.py --config3 > gpurun_out/ev2_probe.txt 2>&1
cut -c1-300 gpurun_out/ev2_bench_c2.json; tail -3 gpurun_out/ev2_bench.err; cat gpurun_out/ev2_smoke.txt | tail -2
