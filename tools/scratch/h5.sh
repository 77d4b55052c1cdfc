timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py -m gpu -x -q 2>&1 | tail -2
run() { # ld epi pf
  VADX_TC_BACKOFF_LD=$1 VADX_TC_BACKOFF_EPI=$2 VADX_LIN_PF=$3 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h5.json 2>gpurun_out/h5.err
  python -c "
import json
for l in open('gpurun_out/h5.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ld$1 epi$2 pf$3', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 64 64 0
run 64 64 1
run 64 64 2
run 64 64 3
run 64 256 0
run 32 128 0
run 128 256 0
run 0 0 0
run 64 64 0
run 64 64 1
