#!/bin/bash
# Evidence run on the GPU box (through gpurun from the repo root); everything lands in gpurun_out/ with the prefix $1.
#   gpurun --timeout 1500 -- 'bash tools/gpu_evidence.sh r2'
# 1. ncu launch list of the bench command (durations; eager steps so that every kernel is a launch of its own)
# 2. DRAM bytes + duration of every plan op of ONE training step (NVTX-renamed kernels) -> tools/ncu_step_traffic.py
# 3. ncu --set full (+ source) of every op that takes >= 2 % of the step, one .ncu-rep, kernels renamed after the plan op
P=${1:-r2}; O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/${P}_bench_under_ncu.log 2>&1
NPVC_NVTX=1 ncu --nvtx --print-nvtx-rename kernel --clock-control none --csv \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --log-file $O/${P}_step_traffic.csv python tools/ncu_target.py 16384 adam > $O/${P}_step_traffic.log 2>&1
python tools/ncu_step_traffic.py $O/${P}_step_traffic.csv $O/${P}_ncu_traffic.json
INC=""; for op in wgrad_g_last dgrad_g_last convT_g3 ln_bwd_g2 wgrad_g2 ln_bwd_e0 ln_bwd_g1 convT_g2 convT_g0 ln_g2 conv_e0 dgrad_g2_odd ln_bwd_e1 conv_e1; do INC="$INC --nvtx-include $op/"; done
NPVC_NVTX=1 ncu --nvtx --print-nvtx-rename kernel $INC --set full --clock-control none --import-source on \
    -o $O/${P}_top_ops python tools/ncu_target.py 16384 > $O/${P}_ncu_top.log 2>&1
tail -2 $O/${P}_ncu_top.log
