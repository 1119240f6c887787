mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log
