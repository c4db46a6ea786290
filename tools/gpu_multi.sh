#!/bin/bash
# N-GPU lines of bench.py (default workload and BASELINE configs 3 and 4), launched the way the driver does
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
RUN="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
$RUN --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
[ -n "$ONLY_DEFAULT" ] || $RUN --workload config4 --steps 2 --warmup 3 > gpurun_out/bench_config4_${TAG}_${N}gpu.json 2>> gpurun_out/bench_${TAG}_${N}gpu.err

for f in bench bench_config4; do echo "== $f"; grep '^{' gpurun_out/${f}_${TAG}_${N}gpu.json | cut -c1-400; done
tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
