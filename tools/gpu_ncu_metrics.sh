# per-launch tensor-pipe / L2 / DRAM metrics of one bench step -> gpurun_out/metrics.csv
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/metrics.csv python tools/profile_step.py > gpurun_out/ncu_metrics.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_metrics.log
