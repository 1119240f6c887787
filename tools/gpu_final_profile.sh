# Round-end evidence of one bench step (config c2) with the final kernels: launch list + DRAM/L2 bytes per launch, the
# traffic file bench.py checks, and ncu --set full pages of EVERY launch of the step (details + raw pages as text).
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/final_plain.log 2>&1 || { tail -5 gpurun_out/final_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_traffic.csv python tools/profile_step.py > gpurun_out/final_ncu_traffic.log 2>&1; echo "traffic exit $?"
python tools/make_traffic_json.py gpurun_out/final_traffic.csv c2 gpurun_out/final_traffic.json
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/final_full python tools/profile_step.py > gpurun_out/final_ncu_full.log 2>&1; echo "full exit $?"
ncu -i /tmp/final_full.ncu-rep --page details > gpurun_out/final_full_details.txt 2>&1
ncu -i /tmp/final_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size > gpurun_out/final_full_raw.csv 2>&1
ls -la /tmp/final_full.ncu-rep gpurun_out/final_full_details.txt gpurun_out/final_full_raw.csv
