#!/bin/bash
# Developer helper, run ON the GPU box (gpurun -- 'bash tools/gpu_ab.sh [SET] [BATCHES]'): A/B of the current build
# against every library under tfhe_gpu_b200/build/ab/ (made by tools/build_variant.sh, or a copy of an older build) with
# tools/abbench.cpp -- same keys, same inputs, outputs compared by checksum.  Results: gpurun_out/abbench_SET.json.
cd "$(dirname "$0")/.."
SET=${1:-std128}; BATCHES=${2:-16384,2048}
AB=tfhe_gpu_b200/build/abbench
[ -x $AB ] || g++ -O2 -o $AB tools/abbench.cpp -I include -I /usr/local/cuda/include -L /usr/local/cuda/lib64 -lcudart -ldl
mkdir -p gpurun_out
timeout 300 $AB --set $SET --batch $BATCHES --reps 5 tfhe_gpu_b200/libtfhe_b200.so $(ls tfhe_gpu_b200/build/ab/*.so 2>/dev/null) \
    > gpurun_out/abbench_$SET.json 2> gpurun_out/abbench_$SET.err
echo "abbench rc $?"
python - <<PY
import json
for l in open('gpurun_out/abbench_$SET.json'):
    d = json.loads(l)
    print(d['set'], d['batch'], d['spec'].split('/')[-1], 'br_ms', d['br_ms_med'], 'total_ms', d['total_ms_med'], 'same bits', d['same_as_first'])
PY
