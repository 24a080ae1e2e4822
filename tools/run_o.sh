mkdir -p gpurun_out
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
(timeout 200 python bench.py --workload cg7_128 --steps 5 --warmup 3 $F) > gpurun_out/r2o_c3.log 2>&1
(timeout 200 python bench.py --workload pcg27_128 --steps 5 --warmup 3 $F) > gpurun_out/r2o_p128.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
(timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:'k_spmv|k_vec' -c 80 --csv --log-file gpurun_out/r2o_c3_ncu.csv python bench.py --workload cg7_128 --steps 1 --warmup 1 --iters 12 $F) > gpurun_out/r2o_c3_ncu.log 2>&1
(timeout 300 ncu --metrics $M --clock-control none --kernel-name regex:'k_spmv|k_vec' -c 80 --csv --log-file gpurun_out/r2o_p128_ncu.csv python bench.py --workload pcg27_128 --steps 1 --warmup 1 --iters 12 $F) > gpurun_out/r2o_p128_ncu.log 2>&1
tail -c 600 gpurun_out/r2o_c3.log
