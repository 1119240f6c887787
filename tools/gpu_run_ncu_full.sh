# usage: bash tools/gpu_run_ncu_full.sh <kernel-regex> <out-name> [launch-skip] [launch-count]
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -s ${3:-0} -c ${4:-3} -f -o gpurun_out/$2 python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full.log
