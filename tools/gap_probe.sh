#!/bin/bash
export LCGB200_LIB=$PWD/tools/variants/s2_t1792_c4.so
for poll in 4 32; do for it in 100 400; do
timeout 300 python bench.py --workload pcg27_256 --steps 3 --iters $it --poll $poll --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; g = d['diagnostics']
print('poll=$poll iters=$it it/s=%.1f ms_per_step=%.2f solve_dev_ms=%.2f prof_dev_ms=%.2f kernel_sum=%.2f spmv=%.4f vec=%.4f launches=%d' % (d['value'], d['ms_per_step'], g['solve_device_ms_per_step'], g['profile_pass_device_ms_per_step'], g['kernel_ms_sum_per_step'], r['avg_launch_ms'], r['vec_kernels']['avg_launch_ms'], d['gpu_launches']))
"
done; done
