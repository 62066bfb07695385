#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
# launch list of ONE training step (the second; the first warms up): ~660 launches after ~700 of warm-up + setup
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python tools/profile_train.py 32 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu.log; wc -l $O/train_launches.csv
