# Development cycle on one GPU box (round 2):  gpurun --timeout 2400 -- 'bash tools/gpu_dev_cycle.sh [tag]'
# tests (all, no -x) -> smoke -> bench -> precision table at the headline shape -> ncu launch list -> per-kernel DRAM / L2
# traffic of one step -> ncu --set full of the non-GEMM kernels and the halo / stem convs (details pages as text).
# Every step logs into gpurun_out/<tag>_*.log; a failing step does not stop the following ones.
TAG=${1:-dev}
mkdir -p gpurun_out; : > gpurun_out/${TAG}_summary.txt
run() { name=$1; shift; limit=$1; shift; timeout "$limit" "$@" > "gpurun_out/${TAG}_$name.log" 2>&1; echo "$name exit $?" >> gpurun_out/${TAG}_summary.txt; }
run pytest_gpu 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -rf
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?" >> gpurun_out/${TAG}_summary.txt
run precision_8x64 900 python tests/gpu_precision_report.py 8 64
run precision_1x90 600 python tests/gpu_precision_report.py 1 90
run bench_arch 600 python tools/bench_arch.py
if [ -z "$SKIP_NCU" ] && timeout 300 python tools/profile_step.py > gpurun_out/${TAG}_plain.log 2>&1; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu launch list exit $?" >> gpurun_out/${TAG}_summary.txt
  M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__cycles_elapsed.avg
  timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_traffic.csv \
    python tools/profile_step.py > gpurun_out/${TAG}_ncu_traffic.log 2>&1
  echo "ncu traffic exit $?" >> gpurun_out/${TAG}_summary.txt
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"apply_kernel|pool|stem_grad|conv_halo|conv_stem|delta_update" -f -o /tmp/${TAG}_full \
    python tools/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
  echo "ncu full exit $?" >> gpurun_out/${TAG}_summary.txt
  ncu -i /tmp/${TAG}_full.ncu-rep --page details > gpurun_out/${TAG}_full_details.txt 2>&1
  ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>&1
  ls -la /tmp/${TAG}_full.ncu-rep >> gpurun_out/${TAG}_summary.txt
  sz=$(stat -c %s /tmp/${TAG}_full.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/${TAG}_full.ncu-rep gpurun_out/; fi
fi
cat gpurun_out/${TAG}_summary.txt; tail -15 gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; tail -n 4 gpurun_out/${TAG}_precision_*.log
