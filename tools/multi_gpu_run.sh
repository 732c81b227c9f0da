#!/bin/bash
# Multi-GPU measurement script (run on an N-GPU box through gpurun --gpus N): the multi-GPU parity tests, strong
# scaling of the headline batch under torchrun, the functional configs at their BASELINE.json global batch, and the
# single-process GPUSetup(numGPUs = N) path.  Usage: tools/multi_gpu_run.sh N TAG [quick|final]
N=${1:-2}; TAG=${2:-r02}; QUICK=${3:-}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() {  # name, nproc, args...
  local name=$1 np=$2; shift 2; port=$((port+1))
  if [ "$np" = 1 ]; then timeout 600 python bench.py "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  else timeout 600 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err; fi
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_$name.json").read().strip().splitlines()[-1]); print("value %.1f e2e %.1f ms/step %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]), d.get("reference_gpu",""))
except Exception as e: print("no json", e)
PY
)"
}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_multi.log
NOREF="--no-cpu-baseline --no-ref-gpu"
if [ "$QUICK" = "final" ]; then   # short confirmation run of the final tree
  run strong_n$N $N --scaling strong --steps 5 --no-cpu-baseline
  run weak_n$N $N --scaling weak --steps 5 $NOREF
  run sp_strong_n$N 1 --single-process --gpus $N --scaling strong --steps 5 $NOREF
  run sp_sign17_n$N 1 --single-process --gpus $N --config sign17 --scaling strong --steps 2 $NOREF
  run func12_n$N $N --config func12 --scaling strong --steps 2 $NOREF
  exit 0
fi
ks="2"; [ "$N" -ge 4 ] && ks="2 4"; [ "$N" -ge 8 ] && ks="2 4 8"
for k in $ks; do
  if [ "$k" = "$N" ]; then run strong_n$k $k --scaling strong --steps 5 --no-cpu-baseline; else run strong_n$k $k --scaling strong --steps 5 $NOREF; fi
done
run sp_strong_n$N 1 --single-process --gpus $N --scaling strong --steps 5 $NOREF
run func12_n$N $N --config func12 --scaling strong --steps 2 $NOREF
run sign17_n$N $N --config sign17 --scaling strong --steps 2 $NOREF
run sp_sign17_n$N 1 --single-process --gpus $N --config sign17 --scaling strong --steps 2 $NOREF
if [ -z "$QUICK" ]; then
  run decomp17_n$N $N --config decomp17 --scaling strong --steps 2 $NOREF
  run sp_func12_n$N 1 --single-process --gpus $N --config func12 --scaling strong --steps 2 $NOREF
  run weak_n$N $N --scaling weak --steps 5 $NOREF
  run ap_n$N $N --config ap --scaling strong --steps 3 $NOREF
fi
