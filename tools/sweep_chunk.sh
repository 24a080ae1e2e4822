#!/bin/bash
# usage: tools/sweep_chunk.sh tag  — quick perf sweep of SpMV chunking on the two single-GPU headline workloads
for cr in 256 512 1024 2048; do
  for wl in pcg27_256 cg7_128; do
    LCGB200_SPMV_CHUNK_ROWS=$cr timeout 300 python bench.py --workload $wl --steps 3 --no-cpu --no-ref-cuda 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('chunk_rows=$cr', d['config']['workload'], 'it/s=%.1f' % d['value'], 'e2e=%.1f' % d['e2e']['value'], 'spmv_ms=%.4f' % r['avg_launch_ms'], 'spmv_frac=%.3f' % r['frac'], 'vec_ms=%.4f' % r['vec_kernels']['avg_launch_ms'], 'iter_frac=%.3f' % r['iteration']['frac_of_peak'])
"
  done
done
