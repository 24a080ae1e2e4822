mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_pytest.log
(timeout 600 python bench.py --steps 5 --warmup 3) > gpurun_out/r2u_bench1.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_bench1.log
(timeout 600 python bench.py --impl reference --steps 5 --warmup 3) > gpurun_out/r2u_bench1_ref.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_bench1_ref.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --iters 40 $F) > gpurun_out/r2u_ncu_launches.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'k_spmv.*EpiDotAlpha' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_r02 python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/r2u_ncu_spmv.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:'k_vec2' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_vec2_r02 python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/r2u_ncu_vec2.log 2>&1
tail -3 gpurun_out/r2u_pytest.log; tail -c 400 gpurun_out/r2u_bench1.log; ls -la gpurun_out/*.ncu-rep
