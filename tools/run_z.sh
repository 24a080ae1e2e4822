mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q -k "compressed or spmv") > gpurun_out/r2z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_pytest.log
(timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-ref-cuda --no-extra-legs) > gpurun_out/r2z_bench.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_bench.log
(timeout 300 python bench.py --workload cg7_128 --steps 5 --warmup 3 --no-cpu --no-ref-cuda --no-extra-legs) > gpurun_out/r2z_bench_cg7.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_bench_cg7.log
F="--no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs --compress"
(timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_spmv_pat$' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_pat_r02 python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/r2z_ncu_pat.log 2>&1
tail -3 gpurun_out/r2z_pytest.log; python - <<'PY'
import json
for f in ("gpurun_out/r2z_bench.log", "gpurun_out/r2z_bench_cg7.log"):
    for l in open(f):
        if l.startswith("{"):
            j = json.loads(l); print(f, j["value"], json.dumps(j["compressed_operator"]))
PY
ls -la gpurun_out/*pat_r02.ncu-rep
