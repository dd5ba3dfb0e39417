#!/bin/bash
# round-2 call AL (1 GPU): final-state records - whole GPU suite, bench with extras and CPU arm, ncu launch list of the bench
# command, --set full captures of the dominant kernels and of the new block-product / small-eigensolver kernels
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $o/r02al_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $o/r02al_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $o/r02al_n1.json 2> $o/r02al_n1.err; tail -2 $o/r02al_n1.err
CMD="python bench.py --no-extras --steps 3 --warmup 3"
timeout 300 $CMD > $o/r02al_plain.json 2> $o/r02al_plain.err || { echo "plain run failed"; tail -5 $o/r02al_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $o/r02al_launches.csv $CMD > $o/r02al_ncu_launches.log 2>&1; tail -2 $o/r02al_ncu_launches.log
for k in symm_panel_kernel j_pass_tma_kernel syrk_streamk_kernel sub_apply2_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $o/r02al_$k $CMD > $o/r02al_ncu_$k.log 2>&1; tail -1 $o/r02al_ncu_$k.log
done
# one spin per launch (the >= 2-GPU shape of the block product) and the small eigensolver: through their unit tools
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sub_apply3_kernel -s 20 -c 1 -f -o $o/r02al_sub_apply3_kernel build/sub_apply_bench 1376 1 50 > $o/r02al_ncu_sub_apply3.log 2>&1; tail -1 $o/r02al_ncu_sub_apply3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:small_eigh_kernel -s 100 -c 1 -f -o $o/r02al_small_eigh_kernel build/small_eigh_test > $o/r02al_ncu_small_eigh.log 2>&1; tail -1 $o/r02al_ncu_small_eigh.log
ls -la $o/r02al_* | head -30
