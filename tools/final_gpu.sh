set -x
P=r2z; O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${P}_gpu_tests.txt; cat $O/${P}_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/${P}_bench.json 2> $O/${P}_bench.err; cut -c1-200 $O/${P}_bench.json
for c in cfg1 cfg3 cfg5 b16; do python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/${P}_bench_$c.json 2> $O/${P}_bench_$c.err; cut -c1-200 $O/${P}_bench_$c.json; done
python bench.py --impl reference --steps 5 --warmup 1 > $O/${P}_bench_reference.json 2> $O/${P}_bench_reference.err; cut -c1-300 $O/${P}_bench_reference.json
python tools/op_profile.py > $O/${P}_ops_per_step.txt 2>&1
bash tools/gpu_evidence.sh $P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
