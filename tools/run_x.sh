mkdir -p gpurun_out
(LCGB200_DEBUG_L2=1 timeout 300 python tools/l2_restore_probe.py) > gpurun_out/r2x_l2_restore.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_l2_restore.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
(timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_spmv$' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_r02 python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/r2x_ncu_spmv.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_vec2$' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_vec2_cg7_r02 python bench.py --workload cg7_128 --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/r2x_ncu_vec2.log 2>&1
(timeout 300 python -m pytest tests -m gpu -x -q -k "SPG or spg or projected or reference_order_real") > gpurun_out/r2x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest.log
grep -v "^\[lcgb200\]" gpurun_out/r2x_l2_restore.log | tail; tail -3 gpurun_out/r2x_pytest.log; ls -la gpurun_out/*r02.ncu-rep
