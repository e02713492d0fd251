cd /root/repo
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r02_topo_n$N.txt
python -c "import os, psutil; print('cpus', os.cpu_count(), 'mem GB', psutil.virtual_memory().total/1e9)" >> gpurun_out/r02_topo_n$N.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/host_bw_probe.py > gpurun_out/r02_host_bw_n$N.json 2> gpurun_out/host_bw_n$N.err
tail -1 gpurun_out/r02_host_bw_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --sequences-total 1024 > gpurun_out/r02_strong_n$N.json 2> gpurun_out/strong_n$N.err
tail -c 600 gpurun_out/r02_strong_n$N.json; tail -2 gpurun_out/strong_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_weak_n$N.json 2> gpurun_out/weak_n$N.err
tail -c 1200 gpurun_out/r02_weak_n$N.json; tail -2 gpurun_out/weak_n$N.err
