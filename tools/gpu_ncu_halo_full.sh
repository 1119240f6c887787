# dram traffic of every conv launch of one bench step (small metric set) + one `--set full` capture with source of
# the dominant launch (Conv3d_2c_3x3 data gradient = last conv_halo_kernel launch of the step)
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__inst_executed_pipe_tc.sum,sm__cycles_elapsed.avg
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -k regex:"conv_halo|conv_umma|conv_stem" --csv --log-file gpurun_out/conv_traffic.csv python tools/profile_step.py > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo -s 29 -c 1 -f -o gpurun_out/r01_halo_dgrad_2c python tools/profile_step.py > gpurun_out/ncu_halo_full.log 2>&1
echo "halo full exit $?"
ncu -i gpurun_out/r01_halo_dgrad_2c.ncu-rep --page details > gpurun_out/r01_halo_dgrad_2c_details.txt 2>&1
