mkdir -p gpurun_out
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
export LCGB200_DEBUG_L2=1
for W in cg7_128 pcg27_128 pcg27_160 pcg27_256; do
  (timeout 200 python bench.py --workload $W --steps 4 --warmup 3 $F) > gpurun_out/r2s_${W}_auto.log 2>&1
done
(LCGB200_L2_PERSIST=1 timeout 200 python bench.py --workload pcg27_160 --steps 4 --warmup 3 $F) > gpurun_out/r2s_pcg27_160_force.log 2>&1
(LCGB200_L2_PERSIST=1 timeout 200 python bench.py --workload pcg27_200 --steps 4 --warmup 3 $F) > gpurun_out/r2s_pcg27_200_force.log 2>&1
for f in gpurun_out/r2s_*.log; do echo $f; grep -o '"value": [0-9.]*' $f | head -1; grep "L2 window" $f | sort | uniq -c | head -2; done
