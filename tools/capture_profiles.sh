#!/bin/bash
# Round-2 evidence, one GPU: bench lines of every workload, the reference arm, the ncu launch list and full captures.
# Run under gpurun from the repo root; everything lands in gpurun_out/ (copy what should be judged into profiles/).
set -x
O=gpurun_out
python bench.py --steps 100 --warmup 5 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_c4_reference.json 2> $O/r02_bench_c4_reference.err
for w in c1 c2 c3 c5; do python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; done
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --synthetic-map > $O/r02_bench_c4_synthetic_map.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/r02_launches_c4.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_view|k_ped_obs|k_footprints|k_dyn_solve|k_dyn_apply|k_view_consts" -s 12 -c 6 -o $O/r02_prof_c4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_ncu_c4.log 2>&1
for w in c1 c3 c5; do
  ncu --set full --clock-control none -k regex:"k_view|k_ped_obs|k_footprints|k_dyn_solve" -s 8 -c 4 -o $O/r02_prof_$w python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_ncu_$w.log 2>&1
done
python tools/loop_rate.py 1024 8192 > $O/r02_loop_rate.txt 2>&1
ls -la $O | tail -30
