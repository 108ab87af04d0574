for f in 0; do
echo "== f64 $f"; python scripts/gpu_p53.py --sites 8192 --chains 16 --burn 8 --sweeps 4 --reps 2 --f64 $f 2>&1 | grep "by kernel\|B1\|B2\|BW\|B0\|nsamples\|root off\|rror" | tail -7
done
python -m pytest tests/test_gpu_invariance.py tests/test_gpu_summary_parity.py tests/test_gpu_toy_exact.py -x -q 2>&1 | tail -3
