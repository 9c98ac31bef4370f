#!/usr/bin/env bash
# Evidence run on one B200 (under gpurun): the default bench line, then - each only after the same command has exited 0
# without ncu - the launch list and one `ncu --set full` capture per dominant kernel, exported as CSV into gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/capture_profiles.sh'
set -uo pipefail
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
python bench.py > $O/r2_bench_default_n1.json 2> $O/r2_bench_default_n1.err || { echo "bench failed"; tail -5 $O/r2_bench_default_n1.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2>/dev/null || echo "reference arm failed"

run_ok() { "$@" > /dev/null 2> $O/last.err || { echo "FAILED without ncu: $*"; tail -3 $O/last.err; return 1; }; }

if run_ok python bench.py --steps 3 --warmup 3 --large-updates 30; then
  $NCU --target-processes application-only --metrics gpu__time_duration.sum -c 30000 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --large-updates 30 > $O/ncu_launches.log 2>&1
fi
if run_ok python bench.py --skip-large --steps 3 --warmup 3; then
  $NCU --set full --import-source on -k regex:ekf_fused_tile_kernel -s 5 -c 2 -f -o $O/r2_prof_fused_tile python bench.py --skip-large --steps 3 --warmup 3 > $O/ncu_tile.log 2>&1
fi
if run_ok python bench.py --only-large --large-updates 60; then
  $NCU --set full --import-source on -k regex:k_large_sweep_mma -s 2 -c 2 -f -o $O/r2_prof_sweep_mma python bench.py --only-large --large-updates 60 > $O/ncu_sweep.log 2>&1
fi
if run_ok python bench.py --only-laser; then
  $NCU --set full --import-source on -k regex:k_circles_scan -s 1 -c 1 -f -o $O/r2_prof_circles python bench.py --only-laser > $O/ncu_circles.log 2>&1
fi
for n in r2_prof_fused_tile r2_prof_sweep_mma r2_prof_circles; do
  if [ -f $O/$n.ncu-rep ]; then
    ncu -i $O/$n.ncu-rep --page raw --csv > $O/${n}_raw.csv 2>/dev/null
    ncu -i $O/$n.ncu-rep --page source --csv --print-source cuda,sass > $O/${n}_source.csv 2>/dev/null
  fi
done
ls -la $O | grep r2_
