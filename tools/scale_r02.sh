# 8-GPU evidence for the end-to-end path: topology, concurrent pinned-H2D ceiling (bound / unbound), bench.py at 8 GPUs
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; } > gpurun_out/r02_topology.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/h2d_microbench.py > gpurun_out/r02_h2d_1.json 2> gpurun_out/r02_h2d_1.err
$TR --nproc-per-node 8 --master-port 29511 tools/h2d_microbench.py > gpurun_out/r02_h2d_8_bind.json 2> gpurun_out/r02_h2d_8_bind.err
$TR --nproc-per-node 8 --master-port 29512 tools/h2d_microbench.py --no-bind > gpurun_out/r02_h2d_8_nobind.json 2> gpurun_out/r02_h2d_8_nobind.err
$TR --nproc-per-node 8 --master-port 29513 tools/h2d_microbench.py --chunks 8 > gpurun_out/r02_h2d_8_bind_chunks8.json 2> gpurun_out/r02_h2d_8_chunks.err
$TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 --no-families > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
VADX_BENCH_NO_NUMA_BIND=1 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 3 --no-families > gpurun_out/r02_bench_8gpu_nobind.json 2> gpurun_out/r02_bench_8gpu_nobind.err
tail -c 400 gpurun_out/r02_bench_8gpu.err
cat gpurun_out/r02_h2d_*.json
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_8gpu.json", "gpurun_out/r02_bench_8gpu_nobind.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["e2e"].get("numa_bind"))
    except Exception as e:
        print(f, "failed", e)
PY
head -30 gpurun_out/r02_topology.txt
