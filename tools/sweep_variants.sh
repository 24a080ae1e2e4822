#!/bin/bash
# perf sweep over the SpMV build variants in tools/variants (LCGB200_LIB override)
for v in default $(ls tools/variants/*.so | xargs -n1 basename | sed 's/\.so$//'); do
  for wl in pcg27_256 cg7_128; do
    if [ "$v" = default ]; then unset LCGB200_LIB; else export LCGB200_LIB=$PWD/tools/variants/$v.so; fi
    timeout 300 python bench.py --workload $wl --steps 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    r = d['roofline']
    print('variant=$v', d['config']['workload'], 'it/s=%.1f' % d['value'], 'e2e=%.1f' % d['e2e']['value'], 'spmv_ms=%.4f' % r['avg_launch_ms'], 'spmv_frac=%.3f' % r['frac'], 'vec_ms=%.4f' % r['vec_kernels']['avg_launch_ms'], 'iter_frac=%.3f' % r['iteration']['frac_of_peak'])
except Exception as e:
    print('variant=$v $wl FAILED', e)
"
  done
done
