timeout 500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py tests/test_gpu_fsmn.py tests/test_gpu_marblenet.py -m gpu -x -q 2>&1 | tail -4
run() { # st lin
  VADX_ST_LOADERS=$1 VADX_LIN_LOADERS=$2 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h2_$1_$2.json 2>gpurun_out/h2_$1_$2.err
  python -c "
import json
for l in open('gpurun_out/h2_$1_$2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('st$1 lin$2', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
"
}
run 8 8; run 16 8; run 16 16; run 16 164; run 8 8; run 16 8; run 16 16; run 16 164
VADX_LIN_LOADERS=16 timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py -m gpu -x -q 2>&1 | tail -2
VADX_LIN_LOADERS=164 timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py -m gpu -x -q 2>&1 | tail -2
