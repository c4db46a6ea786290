#!/bin/bash
# round-2 call I (8 GPUs): the bench line through osa_multi_anneal on the full config 5 (1 048 576 tries),
# multi-device tests on 8 devices, the CLI on all GPUs
TAG=${1:-r2i}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err
grep '^{' gpurun_out/bench_${N}gpu_$TAG.json | cut -c1-160; tail -2 gpurun_out/bench_${N}gpu_$TAG.err | cut -c1-200
python - $TAG $N <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%sgpu_%s.json" % (sys.argv[2], sys.argv[1])) if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["nccl"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["e2e"]["value"])
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -p no:cacheprovider --tb=short --timeout=600 2>&1 | tail -5
