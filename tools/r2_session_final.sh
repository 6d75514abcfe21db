#!/bin/bash
# round-2 profile session (one B200): launch lists and ncu --set full of the kernels that changed this round, the per-level
# probe of the coarse kernels, compute-sanitizer on small grids.  Nothing measured under ncu / the sanitizer is a bench value.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 120 python tools/small_kernel_probe.py > $O/r2_small_kernel_probe.log 2>&1
timeout 120 python tools/profile_cycle.py solve 16385 > $O/r2_plain_solve16385.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_solve_n16385.csv python tools/profile_cycle.py solve 16385 > $O/r2_ncu_solve16385.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_solve_n4097.csv python tools/profile_cycle.py solve 4097 > $O/r2_ncu_solve4097.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cross -c 2 -o $O/r2_full_k_cross python tools/profile_cycle.py solve 16385 > $O/r2_ncu_full_cross.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coarse_cluster|k_coarse_local" -c 3 -o $O/r2_full_coarse python tools/profile_cycle.py solve 4097 > $O/r2_ncu_full_coarse.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_up|k_down" -c 12 -o $O/r2_full_mid_passes python tools/profile_cycle.py solve 4097 > $O/r2_ncu_full_mid.log 2>&1
# compute-sanitizer: memcheck + racecheck on small single-GPU solves (cluster kernel, local coarse kernel, cross pass, slab kernels)
timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_cases.py > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py quick > $O/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/r2_sanitizer_racecheck.log
tail -5 $O/r2_sanitizer_memcheck.log $O/r2_sanitizer_racecheck.log
cat $O/r2_plain_solve16385.log
