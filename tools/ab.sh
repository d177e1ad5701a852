#!/bin/bash
# A/B of library variants on one GPU: tools/ab.sh "<workloads>" <variant.so> ...   ("" = the in-tree library)
# Variants are built with img_env_b200.build.build_variant(out, ["MACRO=1", ...]) and selected through IMGENV_LIB_PATH.
WL=${1:-c4}; shift
for rep in 1 2; do
for v in "$@"; do
  for w in $WL; do
    IMGENV_LIB_PATH=$v python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $w', 'value %.3fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], d['kernel_ms'])"
  done
done
done
