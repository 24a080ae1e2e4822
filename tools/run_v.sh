mkdir -p gpurun_out
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
export LCGB200_DEBUG_L2=1
for W in cg7_128 pcg27_128; do
  for M in 0 1; do
    (LCGB200_VEC2_PDL=$M timeout 150 python bench.py --workload $W --steps 4 --warmup 3 $F) > gpurun_out/r2v_${W}_pdl$M.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_${W}_pdl$M.log
  done
done
(timeout 300 python -m pytest tests -m gpu -x -q -k "real_solvers_match or pinned or full_size") > gpurun_out/r2v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_pytest.log
for f in gpurun_out/r2v_*pdl*.log; do echo $f; grep -o '"value": [0-9.]*' $f | head -1; grep "refused" $f | head -1; tail -1 $f; done; tail -3 gpurun_out/r2v_pytest.log
