#!/bin/bash
# timing experiments on k_sdf_tc2 (results are invalid with a debug flag set; only the kernel time matters)
for dbg in ${T2_DBG_LIST:-0 4 8 16 64 12 124}; do
  SURF_T2_DEBUG=$dbg timeout 90 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mlp-mode 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dbg=$dbg', 'ms_per_step %.1f' % d['ms_per_step'], 'mlp_grad_ms %.1f' % d['roofline']['kernel_ms_per_step']['sdf_mlp_grad'])" || echo "dbg=$dbg failed or timed out"
done
