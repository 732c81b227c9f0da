cd /root/repo
AB=tfhe_gpu_b200/build/abbench; L=tfhe_gpu_b200/libtfhe_b200.so; B=tfhe_gpu_b200/build/ab
timeout 100 $AB --batch 16384,2048,700 --reps 4 $B/base.so $L > gpurun_out/r02y_ab.json 2> gpurun_out/r02y_ab.err; echo "ab rc $?"
timeout 60 $AB --set sign17 --batch 512 --reps 2 $B/base.so $L >> gpurun_out/r02y_ab.json 2>> gpurun_out/r02y_ab.err; echo "ab rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/r02y_ab.json'):
    d=json.loads(l); print(d['set'], d['batch'], d['spec'].split('/')[-1], d['br_ms_med'], d['total_ms_med'], d['same_as_first'])
PY
timeout 150 python -m pytest tests/test_gpu_kernel_matrix.py tests/test_gpu_host_paths.py -q -m gpu -k "persistent or tail_launch" > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02y_pytest.log
