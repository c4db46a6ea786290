#!/bin/bash
# end-of-round check on N GPUs: the multi-device tests, then the default bench line launched the way the driver does
N=${1:-2}; TAG=${2:-multi_final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 2>&1 | tail -3
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
  > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
  > gpurun_out/bench_ref_${TAG}_${N}gpu.json 2> gpurun_out/bench_ref_${TAG}_${N}gpu.err ) 2>&1 | grep real
grep '^{' gpurun_out/bench_ref_${TAG}_${N}gpu.json | cut -c1-200
python - $TAG $N <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s_%sgpu.json" % (sys.argv[1], sys.argv[2])) if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d.get("nccl"))
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"], o["steps"], o["warmup"])
PY
