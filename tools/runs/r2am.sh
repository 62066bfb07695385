#!/bin/bash
# round 2, GPU call O: final state — smoke, whole GPU suite, evidence (launch list + plan-signed summary), full default bench
O=gpurun_out/r2am; mkdir -p $O
python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 1800 python -m pytest tests -m gpu -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/tests.log | cut -c1-300
bash tools/runs/r2i.sh > $O/r2i.log 2>&1; tail -3 $O/r2i.log | cut -c1-400
cp gpurun_out/r2i/ncu_step_summary.json profiles/ncu_step_summary.json
timeout 600 python bench.py --dump-layers $O/layers.json > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-1200 $O/bench.json
