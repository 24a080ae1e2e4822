# row-pattern kernels on bench.py --compress (27-point 256^3 Jacobi-PCG): the box kernel (default for dense box stencils on
# aligned grids) against the general chain kernel (LCGB200_PAT_NO_BOX=1) and the opt-in plane-marching kernel
# (LCGB200_PAT_MARCH=1; LCGB200_PAT_AHEAD = windows loaded ahead, LCGB200_PAT_SEGS = march segments per resident block);
# prints it/s and the SpMV's average launch time.  usage: gpurun -- 'bash tools/run_pat_sweep.sh TAG'
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
run box LCGB200_X=1
run chains LCGB200_PAT_NO_BOX=1
run march LCGB200_PAT_MARCH=1
