timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_firered.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -2
run() { # linpf spread mempf
  VADX_LIN_PF=$1 VADX_LIN_PF_SPREAD=$2 VADX_MEM_PF=$3 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h6.json 2>gpurun_out/h6.err
  python -c "
import json
for l in open('gpurun_out/h6.json'):
    if l.startswith('{'):
        d=json.loads(l); print('linpf$1 spread$2 mempf$3', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 0 0 0
run 1 0 0
run 1 0 1
run 1 0 2
run 1 1 1
run 2 1 1
run 1 1 0
run 1 0 1
run 1 1 1
