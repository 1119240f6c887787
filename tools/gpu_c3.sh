mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_i3d.py -m gpu -q --timeout 600 -p no:cacheprovider -rf > gpurun_out/c3_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/c3_pytest.log
rm -f gpurun_out/halo_ab.txt
bash tools/gpu_halo_ab.sh > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c3_launches.csv python tools/profile_step.py > gpurun_out/c3_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_umma|apply_u8" -c 6 -f -o /tmp/c3_full python tools/profile_step.py > gpurun_out/c3_ncu_full.log 2>&1
ncu -i /tmp/c3_full.ncu-rep --page details > gpurun_out/c3_full_details.txt 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"pool3s1_bwd|pool_s2_bwd" -f -o /tmp/c3_pool python tools/profile_step.py >> gpurun_out/c3_ncu_full.log 2>&1
ncu -i /tmp/c3_pool.ncu-rep --page details > gpurun_out/c3_pool_details.txt 2>&1
cat gpurun_out/halo_ab.txt
