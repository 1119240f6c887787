mkdir -p gpurun_out
FAV_TAP_PROF=1 FAV_BRANCH_STREAMS=0 timeout 300 python tools/profile_step.py 2> gpurun_out/c7_tap_prof.txt > /dev/null; echo "exit $?"
grep "tap prof" gpurun_out/c7_tap_prof.txt | tail -38
