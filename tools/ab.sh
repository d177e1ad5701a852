for v in "" build_variants/v_c3.so build_variants/v_c5.so build_variants/v_c6.so; do
  for w in c4 c1; do
    IMGENV_LIB_PATH=$v python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $w', 'value %.2fM'%(d['value']/1e6), d['kernel_ms'])"
  done
done
