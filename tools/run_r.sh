mkdir -p gpurun_out
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
for W in cg7_128 pcg27_128 pcg27_160 pcg27_200 pcg27_256; do
  for M in 0 1; do
    (LCGB200_L2_PERSIST=$M LCGB200_DEBUG_L2=1 timeout 200 python bench.py --workload $W --steps 4 --warmup 3 $F) > gpurun_out/r2r_${W}_l2_$M.log 2>&1
  done
done
for f in gpurun_out/r2r_*.log; do echo $f; grep -o '"value": [0-9.]*' $f | head -1; grep "L2 window" $f | head -1; done
