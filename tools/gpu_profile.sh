#!/bin/bash
# ncu evidence: launch list of one step + full capture of the hot kernels.
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --part step > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?" >> gpurun_out/ncu_step.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -f -o gpurun_out/prof_hot python tools/profile_step.py --part hot > gpurun_out/ncu_hot.log 2>&1
echo "ncu hot exit $?" >> gpurun_out/ncu_hot.log
tail -n 3 gpurun_out/ncu_step.log; tail -n 3 gpurun_out/ncu_hot.log; ls -la gpurun_out
