set -x
cap() { # family regex count [env]
  fam=$1; rx=$2; n=$3
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_$fam.csv python tools/family_step.py $fam eager > gpurun_out/r02_launches_$fam.log 2>&1
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$rx" -c $n -f -o gpurun_out/r02_prof_$fam python tools/family_step.py $fam eager > gpurun_out/r02_prof_$fam.log 2>&1
  tail -2 gpurun_out/r02_prof_$fam.log
  # summarise ON the box: gpurun only brings back 64 MiB, the .ncu-rep files are 10-35 MB each
  python tools/summarize_ncu.py full gpurun_out/r02_prof_$fam.ncu-rep > gpurun_out/r02_top_kernels_$fam.txt
  python tools/summarize_ncu.py launches gpurun_out/r02_launches_$fam.csv > gpurun_out/r02_launches_$fam.txt
  rm -f gpurun_out/r02_prof_$fam.ncu-rep gpurun_out/r02_launches_$fam.csv
}
cap firered 'fc2_memory_stages_kernel|linear_tc_kernel|stft_power_tc_kernel' 8
cap fsmn 'linear_tc_kernel|stft_power_tc_kernel|fsmn_memory|lfr_cmvn' 8
cap marblenet 'stft_power_tc_kernel|depthwise_conv1d|linear_tc_kernel' 8
cap silero 'linear_tc_kernel|gemm_f32_kernel|lstm_cell|reflect' 8
export VADX_BENCH_DFSMN_PAIRS=128
cap dfsmn_aec 'lstm|gemm_f32_kernel|layernorm|permute4' 10
unset VADX_BENCH_DFSMN_PAIRS
cap firered_stream 'fsmn_memory|linear_tc_kernel' 6
ls -la gpurun_out/r02_*; du -sh gpurun_out
