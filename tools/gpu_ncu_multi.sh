# usage: bash tools/gpu_ncu_multi.sh  -> ncu --set full captures of the pool / stem / sat kernels of one bench step
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
cap() {  # regex out skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$1" -s $3 -c $4 -f -o gpurun_out/$2 python tools/profile_step.py > gpurun_out/ncu_$2.log 2>&1
  echo "$2 exit $?"
}
cap maxpool_bwd prof_poolbwd 0 20
cap maxpool_fwd prof_poolfwd 0 20
cap "conv_umma" prof_stem 0 2
cap "stem_sat|stem_class|head_logits|loss_kernel|apply" prof_misc 0 8
