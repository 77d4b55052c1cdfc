timeout 500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py tests/test_gpu_fsmn.py tests/test_gpu_marblenet.py -m gpu -x -q 2>&1 | tail -4
run() { # st_lw st_opt lin_lw no256
  if [ "$4" = "1" ]; then export VADX_LIN_NO_LDG256=1; else unset VADX_LIN_NO_LDG256; fi
  VADX_ST_LOADERS=$1 VADX_ST_OPT=$2 VADX_LIN_LOADERS=$3 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h3.json 2>gpurun_out/h3.err
  python -c "
import json
for l in open('gpurun_out/h3.json'):
    if l.startswith('{'):
        d=json.loads(l); print('st$1 opt$2 lin$3 no256=$4', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 8 0 8 1
run 8 1 8 1
run 8 2 8 1
run 8 3 8 0
run 16 3 8 0
run 16 3 16 0
run 8 3 16 0
run 8 3 16 1
run 8 0 8 1
run 8 3 16 0
