mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/plain.log
