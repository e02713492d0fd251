"""Host <-> device copy ceiling of a multi-GPU box, no kernels involved (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/host_bw_probe.py

Every rank copies pinned host memory to its GPU (and back) at the same time as all the others, in the proportion of
bench.py's end-to-end step (3360 B of detections in, ~1380 B of results out per frame).  The aggregate H2D rate divided
by 3360 B is the ceiling of the `e2e` metric on this box, whatever the kernels do."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world)
GB = 1024 ** 3
n_in, n_out = 8 * GB // 4, int(8 * GB * 1380 / 3360) // 4
h_in = torch.empty(n_in, dtype=torch.float32, pin_memory=True)
h_out = torch.empty(n_out, dtype=torch.float32, pin_memory=True)
h_in.fill_(1.0)
d_in = torch.empty(n_in, dtype=torch.float32, device="cuda")
d_out = torch.zeros(n_out, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        best = min(best, el)
    return best


def h2d():
    chunk = n_in // 16
    for c in range(16):
        with torch.cuda.stream(s1):
            d_in[c * chunk:(c + 1) * chunk].copy_(h_in[c * chunk:(c + 1) * chunk], non_blocking=True)


def both():
    h2d()
    chunk = n_out // 16
    for c in range(16):
        with torch.cuda.stream(s2):
            h_out[c * chunk:(c + 1) * chunk].copy_(d_out[c * chunk:(c + 1) * chunk], non_blocking=True)


t_in = timed(h2d)
t_both = timed(both)
if rank == 0:
    agg_in = world * n_in * 4 / t_in / 1e9
    agg_both_in = world * n_in * 4 / t_both / 1e9
    res = {"gpus": world, "host_cpus": os.cpu_count(),
           "h2d_only_aggregate_GBps": agg_in, "h2d_only_per_gpu_GBps": agg_in / world,
           "h2d_with_concurrent_d2h_aggregate_GBps": agg_both_in, "d2h_concurrent_aggregate_GBps": world * n_out * 4 / t_both / 1e9,
           "e2e_ceiling_frames_per_s": agg_both_in * 1e9 / 3360.0,
           "note": "pure cudaMemcpyAsync from/to pinned memory on all ranks at once (max over ranks); no kernel runs"}
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
