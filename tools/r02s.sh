#!/bin/bash
# round-2 call S: ncu launch list of the bench command and --set full captures of the dominant kernels
cd "$(dirname "$0")/.."
o=gpurun_out
CMD="python bench.py --no-extras --steps 3 --warmup 3"
timeout 300 $CMD > $o/r02s_plain.json 2> $o/r02s_plain.err || { echo "plain run failed"; tail -5 $o/r02s_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $o/r02s_launches.csv $CMD > $o/r02s_ncu_launches.log 2>&1; tail -2 $o/r02s_ncu_launches.log
for k in syrk_streamk_kernel j_pass_tma_kernel symm_panel_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $o/r02s_$k $CMD > $o/r02s_ncu_$k.log 2>&1; tail -2 $o/r02s_ncu_$k.log
done
ls -la $o/r02s_*
