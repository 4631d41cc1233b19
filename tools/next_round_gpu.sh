#!/bin/bash
# First GPU call of the next round (run through gpurun from the repo root, ~3 GPU-minutes):
#   gpurun --timeout 900 -- 'bash tools/next_round_gpu.sh'
# 1. parity suite + bench line with the CTA-pair form on by default (the round-1 bench line predates it)
# 2. the experiments that were written without a GPU, each against a default engine on the same inputs:
#    wgrad pair kernel, two-stream half-batches, trimmed last k-block, 128-column N tiles
# 3. cfg3 / cfg5 timings, launch list and ncu --set full of the pair kernel (G3 forward)
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2a_gpu_tests.txt 2>&1; tail -3 $O/r2a_gpu_tests.txt
NPVC_TEST_EXPERIMENTS=1 python -m pytest tests -m gpu -q -k experimental > $O/r2a_gpu_experiments.txt 2>&1; tail -15 $O/r2a_gpu_experiments.txt
python bench.py --steps 20 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; cut -c1-400 $O/r2a_bench.json
for sw in NPVC_WGRAD_PAIR=1 NPVC_WGRAD_PAIR=2 NPVC_STREAMS=2 NPVC_PAIR_TRIM=1 NPVC_BN_CAP=128 "NPVC_BN_CAP=128 NPVC_BN_CAP_K=1024"; do
  f=$O/r2a_switch_$(echo $sw | tr ' =' '__').txt
  timeout 60 python tools/pair_check.py 16384 default $sw > $f 2>&1; echo "rc=$?" >> $f
  echo "== $sw"; grep -E "pair vs base|ms/pass|rc=" $f
done
timeout 120 python tools/infer_time.py > $O/r2a_infer_cfg3.txt 2>&1; cat $O/r2a_infer_cfg3.txt
timeout 120 python tools/cfg5_time.py > $O/r2a_cfg5.txt 2>&1; cat $O/r2a_cfg5.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2a_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_fwd_kernel_t -c 12 -o $O/r2a_fwd python tools/ncu_target.py > $O/r2a_ncu_fwd.log 2>&1
