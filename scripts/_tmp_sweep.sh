for ls in 0 4 5 10 20; do for w in 0 16; do
echo "== lockstep $ls warps $w"; python scripts/gpu_p53.py --sites 8192 --chains 16 --burn 8 --sweeps 4 --reps 2 --lockstep $ls --warps $w 2>&1 | grep "by kernel" | tail -1
done; done
