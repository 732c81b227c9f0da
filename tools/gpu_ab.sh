cd /root/repo
AB=tfhe_gpu_b200/build/abbench; L=tfhe_gpu_b200/libtfhe_b200.so; B=tfhe_gpu_b200/build/ab
timeout 200 $AB --batch 16384,2048 --reps 5 $B/base.so $L $L:persistent=0 $B/t2.so:persistent=0 $B/t4.so:persistent=0 $B/t2.so $B/t4.so > gpurun_out/r02u_ab2.json 2> gpurun_out/r02u_ab2.err; echo "ab rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/r02u_ab2.json'):
    d=json.loads(l); print(d['batch'], d['spec'].split('/')[-1], d['br_ms_med'], d['total_ms_med'], d['same_as_first'])
PY
tail -3 gpurun_out/r02u_ab2.err
