timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/op_profile.py 2>&1 | sort -k1,1 > /tmp/new.txt
NPVC_UMMA_DUAL=0 python tools/op_profile.py 2>&1 | sort -k1,1 > /tmp/old.txt
join /tmp/new.txt /tmp/old.txt | awk '{d=$2-$10; if (d>0.003||d<-0.003) printf "%-16s dual %.4f single %.4f  diff %+.4f\n",$1,$2,$10,d}'
grep total /tmp/new.txt /tmp/old.txt
python bench.py --no-cpu-baseline 2>/dev/null | cut -c150-200
NPVC_UMMA_DUAL=0 python bench.py --no-cpu-baseline 2>/dev/null | cut -c150-200
