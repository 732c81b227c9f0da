cd /root/repo
AB=tfhe_gpu_b200/build/abbench; L=tfhe_gpu_b200/libtfhe_b200.so; B=tfhe_gpu_b200/build/ab
timeout 200 $AB --batch 16384,2048 --reps 5 $B/base.so $L $B/i7.so $B/m2.so > gpurun_out/r02u_ab4.json 2> gpurun_out/r02u_ab4.err; echo "ab rc $?"
timeout 200 $AB --set func12 --batch 8192,1024 --reps 3 $B/base.so $L >> gpurun_out/r02u_ab4.json 2>> gpurun_out/r02u_ab4.err; echo "ab rc $?"
timeout 200 $AB --set sign17 --batch 4096,512 --reps 3 $B/base.so $L >> gpurun_out/r02u_ab4.json 2>> gpurun_out/r02u_ab4.err; echo "ab rc $?"
python - <<'PY'
import json
for l in open('gpurun_out/r02u_ab4.json'):
    d=json.loads(l); print(d['set'], d['batch'], d['spec'].split('/')[-1], d['br_ms_med'], d['total_ms_med'], d['same_as_first'])
PY
tail -3 gpurun_out/r02u_ab4.err
