# Round-2 ncu evidence per workload (run under gpurun on one B200): a launch list of one step (gpu__time_duration, cold-cache,
# serialised: compare shares) and --set full captures: EVERY launch of the family's dominant kernel in one step (so that the
# mean dram bytes per launch is comparable with bench.py's mean algorithmic bytes per launch) plus a few of the next kernels.
# Summaries are made ON the box (gpurun brings back 64 MiB; the .ncu-rep files are 10-35 MB each).
set -x
mkdir -p gpurun_out
cap() { # family  dominant-regex  count  others-regex  count
  fam=$1
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_$fam.csv python tools/family_step.py $fam eager > gpurun_out/r02_launches_$fam.log 2>&1
  python tools/summarize_ncu.py launches gpurun_out/r02_launches_$fam.csv > gpurun_out/r02_launches_$fam.txt
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" -c $3 -f -o gpurun_out/r02_prof_${fam}_a python tools/family_step.py $fam eager > gpurun_out/r02_prof_$fam.log 2>&1
  python tools/summarize_ncu.py full gpurun_out/r02_prof_${fam}_a.ncu-rep > gpurun_out/r02_top_kernels_$fam.txt
  if [ -n "$4" ]; then
    ncu --set full --clock-control none --profile-from-start off -k regex:"$4" -c $5 -f -o gpurun_out/r02_prof_${fam}_b python tools/family_step.py $fam eager >> gpurun_out/r02_prof_$fam.log 2>&1
    python tools/summarize_ncu.py full gpurun_out/r02_prof_${fam}_b.ncu-rep | grep -v "^#" >> gpurun_out/r02_top_kernels_$fam.txt
  fi
  rm -f gpurun_out/r02_prof_${fam}_*.ncu-rep gpurun_out/r02_launches_$fam.csv
  tail -1 gpurun_out/r02_prof_$fam.log
}
for fam in "$@"; do
  case $fam in
    firered)        cap firered 'fc2_memory_stages_kernel' 8 'linear_tc_kernel|stft_power_tc_kernel' 11 ;;
    fsmn)           cap fsmn 'linear_tc_kernel' 12 'stft_power_tc_kernel|lfr_cmvn|fsmn_memory|gather_windows' 5 ;;
    marblenet)      cap marblenet 'stft_power_tc_kernel' 1 'depthwise_conv1d_reg_kernel|linear_tc_kernel' 10 ;;
    silero)         cap silero 'linear_tc_kernel' 12 'silero_lstm_windows_kernel|stft_mag_compact|reflect_window' 3 ;;
    dfsmn_aec)      cap dfsmn_aec 'linear_tc_kernel' 24 'lstm_rec_kernel|cfb_front_kernel|layernorm_perm_kernel' 8 ;;
    firered_stream) cap firered_stream 'fsmn_memory' 8 'linear_tc_kernel' 6 ;;
  esac
done
ls -la gpurun_out/r02_*; du -sh gpurun_out
