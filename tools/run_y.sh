mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2y_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_pytest.log
F="--no-ref-cuda --no-compressed-leg --no-extra-legs"
for W in case10k_cg case10k_pcg; do
  (timeout 200 python bench.py --workload $W --steps 5 --warmup 3 $F) > gpurun_out/r2y_$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_$W.log
done
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2y_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_smoke.log
tail -3 gpurun_out/r2y_pytest.log; for W in case10k_cg case10k_pcg; do grep -o '"value": [0-9.]*' gpurun_out/r2y_$W.log | head -1; done; tail -2 gpurun_out/r2y_smoke.log
