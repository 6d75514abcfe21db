#!/bin/bash
# round-2 closing session (one B200): GPU test suite, smoke, default bench line, the latency probe DESIGN.md cites, launch list of
# the folded F-cycle and ncu --set full of Pass A's prolong-in form.  Nothing measured under ncu is a bench value.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r2f_pytest.log 2>&1; tail -3 $O/r2f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; tail -1 $O/r2f_smoke.log
timeout 900 python bench.py > $O/r2f_bench.json 2> $O/r2f_bench.err; tail -c 300 $O/r2f_bench.err
timeout 300 python tools/latency_probe.py > $O/r2_latency_probe_cluster.log 2>&1; tail -1 $O/r2_latency_probe_cluster.log
timeout 120 python tools/profile_cycle.py fcycle 16385 > $O/r2f_plain_fcycle.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_fcycle_n16385.csv python tools/profile_cycle.py fcycle 16385 > $O/r2f_ncu_fcycle.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_down<2, 3, 5" -c 12 -o $O/r2f_full_k_down_prolong python tools/profile_cycle.py fcycle 16385 > $O/r2f_ncu_full_pin.log 2>&1
cat $O/r2f_plain_fcycle.log
