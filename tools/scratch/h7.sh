run() { # at lw
  VADX_LIN_PF_AT=$1 VADX_LIN_LOADERS=$2 python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h7.json 2>gpurun_out/h7.err
  python -c "
import json
for l in open('gpurun_out/h7.json'):
    if l.startswith('{'):
        d=json.loads(l); print('at$1 lw$2', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 0 16
run 1 16
run 2 16
run 0 8
run 2 8
run 0 16
run 1 16
run 2 16
