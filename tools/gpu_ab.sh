cd /root/repo
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --scaling strong --steps 5 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-pageable --no-imad-peak > gpurun_out/r02x_8gpu_strong_n8.json 2> gpurun_out/r02x_8gpu_strong_n8.err; echo "bench rc $?"
tail -c 600 gpurun_out/r02x_8gpu_strong_n8.json | head -c 600; tail -2 gpurun_out/r02x_8gpu_strong_n8.err
