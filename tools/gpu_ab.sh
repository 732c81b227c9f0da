cd /root/repo
timeout 400 python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_full_size.py -q -m gpu > gpurun_out/r02x_pytest_host.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02x_pytest_host.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02x_bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pageable',d['e2e_pageable']['value'],'frac',d['roofline']['frac'])
PY
