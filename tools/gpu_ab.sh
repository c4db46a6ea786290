#!/bin/bash
# A/B of the sweep kernel on the GPU: parity tests with the in-tree library, then tools/probe.py
# dense for every library given as LABEL=PATH (the in-tree build is always run last as "tree").
TAG=${1:-ab}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
for kv in "$@"; do
  label=${kv%%=*}; path=${kv#*=}
  OSA_LIB_PATH=$path timeout 600 python tools/probe.py dense > gpurun_out/probe_${TAG}_$label.log 2>&1
  echo "== $label"; cat gpurun_out/probe_${TAG}_$label.log | cut -c1-700
done
timeout 600 python tools/probe.py dense > gpurun_out/probe_${TAG}_tree.log 2>&1
echo "== tree"; cat gpurun_out/probe_${TAG}_tree.log | cut -c1-700
