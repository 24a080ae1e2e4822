#!/bin/bash
# chunk sweep for the dictionary-compressed operator (LCGB200_CSR_COMPRESS)
for cr in 256; do
  LCGB200_SPMV_CHUNK_ROWS=$cr timeout 300 python bench.py --compress --steps 3 --no-cpu --no-ref-cuda 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('dict chunk_rows=$cr it/s=%.1f e2e=%.1f spmv_ms=%.4f frac=%.3f vec_ms=%.4f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], r['frac'], r['vec_kernels']['avg_launch_ms']))"
done
