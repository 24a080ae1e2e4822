# round-end check on one B200: the -m gpu suite, smoke(), the default bench line, the ncu launch list
# usage: gpurun -- 'bash tools/run_final.sh TAG [quick]'
mkdir -p gpurun_out
T=${1:-r2f}
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
(timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
tail -3 gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_smoke.log
if [ "$2" != "quick" ]; then
(timeout -s KILL 600 python bench.py --steps 5 --warmup 3) > gpurun_out/${T}_bench1.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench1.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs"
(timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 1 --iters 40 $F) > gpurun_out/${T}_ncu_launches.log 2>&1
tail -c 300 gpurun_out/${T}_bench1.log
fi
