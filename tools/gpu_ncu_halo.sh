mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo -s 0 -c 1 -f -o gpurun_out/prof_halo_fwd python tools/profile_step.py > gpurun_out/ncu_full1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_halo -s 29 -c 1 -f -o gpurun_out/prof_halo_dgrad python tools/profile_step.py > gpurun_out/ncu_full2.log 2>&1
echo "ncu exit $?"
