cd /root/repo
N=${1:-2}
timeout 300 python -m pytest tests/test_golden_gpu.py tests/test_evaluate.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/host_bw_probe.py > gpurun_out/r02_host_bw_n$N.json 2> gpurun_out/host_bw_n$N.err
tail -2 gpurun_out/r02_host_bw_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --sequences-total 1024 > gpurun_out/r02_strong_n$N.json 2> gpurun_out/strong_n$N.err
tail -c 1500 gpurun_out/r02_strong_n$N.json; tail -3 gpurun_out/strong_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_weak_n$N.json 2> gpurun_out/weak_n$N.err
tail -c 2500 gpurun_out/r02_weak_n$N.json; tail -3 gpurun_out/weak_n$N.err
