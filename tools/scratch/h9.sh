run() { # allcols
  if [ "$1" = "1" ]; then export VADX_TC_ALLCOLS=1; else unset VADX_TC_ALLCOLS; fi
  python bench.py --no-families --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/h9.json 2>gpurun_out/h9.err
  python -c "
import json
for l in open('gpurun_out/h9.json'):
    if l.startswith('{'):
        d=json.loads(l); print('allcols$1', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms_per_step'].items() if v>0.01})
"
}
run 1; run 0; run 1; run 0; run 1; run 0
