#!/bin/bash
# Development helper: time the step kernel of every library variant under _variants/ (built by hand with
# different -D flags) on C2 and C3.  Usage (on the GPU box): bash scripts/variant_sweep.sh [workloads...]
WL=${@:-"c2 c3"}
mkdir -p gpurun_out
cp d3p_b200/_lib/libd3p_b200.so /tmp/lib_main.so
for v in _variants/lib_*.so; do
  cp $v d3p_b200/_lib/libd3p_b200.so
  for w in $WL; do
    python bench.py --workload $w --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-other-workloads --no-ncu-side-run 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v','$w','value=%.4g'%d['value'],'ms_per_step=%.4f'%d['ms_per_step'],'kernel_ms=%.4f'%d['roofline']['kernel_ms'],'frac=%.4f'%d['roofline']['frac'])"
  done
done | tee gpurun_out/variant_sweep.log
cp /tmp/lib_main.so d3p_b200/_lib/libd3p_b200.so
