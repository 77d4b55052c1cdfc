for rg in 13 26; do VADX_MEM_RG=$rg timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_firered.py tests/test_gpu_firered_stream.py -m gpu -x -q 2>&1 | tail -1; done
run() {
  VADX_MEM_RG=$1 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h10.json 2>gpurun_out/h10.err
  python -c "
import json
for l in open('gpurun_out/h10.json'):
    if l.startswith('{'):
        d=json.loads(l); print('rg$1', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 7; run 13; run 26; run 7; run 13; run 26
