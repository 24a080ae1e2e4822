mkdir -p gpurun_out
T=${1:-r2zb}
(timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "compressed or spmv") > gpurun_out/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
if grep -q "rc=0" gpurun_out/${T}_pytest.log; then
(timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-ref-cuda --no-extra-legs) > gpurun_out/${T}_bench.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench.log
(timeout -s KILL 300 python bench.py --workload cg7_128 --steps 5 --warmup 3 --no-cpu --no-ref-cuda --no-extra-legs) > gpurun_out/${T}_bench_cg7.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_bench_cg7.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs --compress"
(timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_spmv_pat' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_pat_${T} python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/${T}_ncu_pat.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/${T}_bench.log", "gpurun_out/${T}_bench_cg7.log"):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l); print(f, j["value"], json.dumps(j["compressed_operator"]))
PY
ls -la gpurun_out/*${T}.ncu-rep
fi
