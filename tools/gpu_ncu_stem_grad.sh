# `ncu --set full` of the stem gradient kernel (and the stem forward) of one bench step; details pages as text
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"stem_grad_kernel|conv_stem_kernel" -c 2 -f -o gpurun_out/r01_stem python tools/profile_step.py > gpurun_out/ncu_stem.log 2>&1
echo "stem full exit $?"
ncu -i gpurun_out/r01_stem.ncu-rep --page details > gpurun_out/r01_stem_details.txt 2>&1
tail -3 gpurun_out/ncu_stem.log
