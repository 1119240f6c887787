# First GPU call of a round, everything in one box: run as
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_open.sh'
# 1. the whole -m gpu suite, 2. smoke(), 3. the bench line, 4. the precision table at the BASELINE shapes
# (tests/gpu_precision_report.py), 5. the plain-PyTorch-on-the-same-GPU timing, 6. the torch-stack
# step times, 7. the ncu launch list of one step (only after the same command exited 0 without ncu).
# Every step writes its own log under gpurun_out/ and a line into gpurun_out/summary.txt; a failing step does not
# stop the following ones.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { name=$1; shift; limit=$1; shift; timeout "$limit" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; }
run pytest_gpu 1200 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
run precision_8x64 900 python tests/gpu_precision_report.py 8 64
run precision_1x90 600 python tests/gpu_precision_report.py 1 90
run precision_1x16 600 python tests/gpu_precision_report.py 1 16
for a in "i3d 8 64" "r3d_18 16 16" "r2plus1d_18 16 16"; do
  n=$(echo $a | tr ' ' '_')
  run torch_ref_$n 600 python tests/gpu_torch_reference_timing.py $a
  run torch_ref_tf32_$n 600 python tests/gpu_torch_reference_timing.py $a --tf32
done
run bench_arch 600 python tools/bench_arch.py
if timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
  echo "ncu launch list exit $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -n 5 gpurun_out/precision_*.log; tail -n 1 gpurun_out/torch_ref_*.log
