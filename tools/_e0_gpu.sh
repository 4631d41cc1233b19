python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/op_profile.py 2>&1 | grep -E "total|conv_e0|ln_bwd_e0"
python tools/op_profile.py 16 2>&1 | grep -E "total|conv_e0|ln_bwd_e0"
python bench.py 2>/dev/null | cut -c1-330
