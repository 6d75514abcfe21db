#!/bin/bash
# round-2 GPU session 1 (one B200): parity suite, bench with the new legs, launch lists and ncu of the small-level kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest1.log
python bench.py --steps 5 --warmup 3 > $O/r2_bench1.log 2> $O/r2_bench1.err; echo "bench rc=$?" >> $O/r2_bench1.err
python tools/profile_cycle.py cycle 4097 > $O/r2_plain_cycle4097.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2_launches_v_n4097.csv python tools/profile_cycle.py cycle 4097 > $O/r2_ncu_v4097.log 2>&1
python tools/profile_cycle.py wcycle 1025 > $O/r2_plain_wcycle1025.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_w_n1025.csv python tools/profile_cycle.py wcycle 1025 > $O/r2_ncu_w1025.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_vcycle_small2 -c 3 -o $O/r2_small2_w python tools/profile_cycle.py wcycle 1025 > $O/r2_ncu_small2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_down|k_up" -c 14 -o $O/r2_mid_levels python tools/profile_cycle.py cycle 1025 > $O/r2_ncu_mid.log 2>&1
tail -3 $O/r2_pytest1.log; cat $O/r2_bench1.err | tail -5; head -c 3000 $O/r2_bench1.log
