mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"pool3s1" -c 4 -f -o gpurun_out/prof_pool3 python tools/profile_step.py > gpurun_out/ncu_pool3.log 2>&1
echo "exit $?"
ncu -i gpurun_out/prof_pool3.ncu-rep --page details 2>&1 | grep -E "pool3s1|Duration|Throughput|Issue Slots Busy|Executed Ipc|No Eligible|Eligible Warps|Achieved Occupancy|Theoretical Occ|Registers Per|Stall|stall|L1/TEX Hit|Bank conflict|Mem Busy|Max Bandwidth|Mem Pipes|Shared Memory|warp cycles per" > gpurun_out/pool3_details.txt
rm -f gpurun_out/prof_pool3.ncu-rep
