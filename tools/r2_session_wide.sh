#!/bin/bash
# round-2, after the wide cross-cycle pass (strips of 128 columns): GPU test suite, launch lists and ncu --set full of k_cross.
# Nothing measured under ncu is a bench value.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2w_pytest.log 2>&1; tail -3 $O/r2w_pytest.log
timeout 120 python tools/profile_cycle.py solve 16385 > $O/r2w_plain_solve16385.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2w_launches_solve_n16385.csv python tools/profile_cycle.py solve 16385 > $O/r2w_ncu_solve16385.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2w_launches_solve_n4097.csv python tools/profile_cycle.py solve 4097 > $O/r2w_ncu_solve4097.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cross -c 2 -o $O/r2w_full_k_cross python tools/profile_cycle.py solve 16385 > $O/r2w_ncu_full_cross.log 2>&1
cat $O/r2w_plain_solve16385.log
