# row-pattern kernels on bench.py --compress (27-point 256^3 Jacobi-PCG): the box kernel (default for dense box stencils) against
# the chain kernel (LCGB200_PAT_NO_BOX=1) and the opt-in plane-marching kernel; prints it/s and the SpMV's average launch time
mkdir -p gpurun_out
T=${1:-sweep}
F="--steps 3 --warmup 3 --no-cpu --no-ref-cuda --no-compressed-leg --no-extra-legs --compress"
run() {  # name, env...
  name=$1; shift
  (env "$@" timeout -s KILL 200 python bench.py $F) > gpurun_out/${T}_$name.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/${T}_$name.log"):
    if l.startswith("{"):
        j = json.loads(l); print("$name", round(j["value"], 1), "it/s  spmv_ms", round(j["roofline"]["avg_launch_ms"], 4), "parity", j["parity"]["ok"] if j.get("parity") else None, j["parity"].get("rel_l2") if j.get("parity") else None)
PY
}
(timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "compressed or march or spmv") 2>&1 | tail -4
run box2 LCGB200_PAT_BOX_BLOCKS=2
run box3 LCGB200_PAT_BOX_BLOCKS=3
(timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name regex:'^k_spmv_pat' --launch-skip 20 --launch-count 1 -f -o gpurun_out/ncu_spmv_pat_${T} python bench.py --steps 1 --warmup 1 --iters 40 $F) > gpurun_out/${T}_ncu_pat.log 2>&1
ls -la gpurun_out/*${T}.ncu-rep
