mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k reference_order) > gpurun_out/r2q_pytest_exact.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest_exact.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
for W in cg7_128 pcg27_128; do
  (timeout 200 python bench.py --workload $W --steps 5 --warmup 3 $F) > gpurun_out/r2q_${W}_base.log 2>&1
  (LCGB200_L2_PERSIST=1 LCGB200_DEBUG_L2=1 timeout 200 python bench.py --workload $W --steps 5 --warmup 3 $F) > gpurun_out/r2q_${W}_l2.log 2>&1
done
tail -3 gpurun_out/r2q_pytest_exact.log
for f in gpurun_out/r2q_*_base.log gpurun_out/r2q_*_l2.log; do echo $f; grep -o '"value": [0-9.]*' $f | head -1; grep "L2 window" $f | head -1; done
