# round-end check on one B200: the -m gpu suite, smoke(), the default bench line and the reference arm, the ncu launch list
mkdir -p gpurun_out
T=${1:-r2f}
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
(timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
(timeout -s KILL 600 python bench.py --steps 5 --warmup 3) > gpurun_out/${T}_bench1.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench1.log

F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
(timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --iters 40 $F) > gpurun_out/${T}_ncu_launches.log 2>&1
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log; tail -c 300 gpurun_out/${T}_bench1.log; 
F2="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs --compress"
(timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_spmv_pat' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_pat_${T} python bench.py --steps 1 --warmup 1 --iters 40 $F2) > gpurun_out/${T}_ncu_pat.log 2>&1
(timeout -s KILL 200 python bench.py --workload cg7_128 --compress --steps 5 --warmup 3 --no-cpu --no-ref-cuda --no-extra-legs) > gpurun_out/${T}_cg7_compressed.log 2>&1
grep -o '"value": [0-9.]*' gpurun_out/${T}_cg7_compressed.log | head -1
