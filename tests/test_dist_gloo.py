"""Multi-process plumbing (world_size 2, gloo, CPU): sequence sharding, counter all-reduce, result
gather -- the only collectives of the path.  The per-rank work is done by
the oracle here (no GPU in this container); on the B200 box the same helpers run over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import pam_b200  # noqa: F401
    from pam_b200 import dist as pdist, synth
    from oracle import generic
    r, lr, w = pdist.init("gloo")
    assert (r, w) == (rank, world)
    total, T = 5, 25
    mine = pdist.shard_sequences(total, r, w)
    counts = torch.zeros((3, T), dtype=torch.int32)            # fixed-stride block: ceil(5/2) sequences
    ids = torch.full((3, T, 8), -1, dtype=torch.int32)
    counters = torch.zeros(2, dtype=torch.int64)
    for k, s in enumerate(mine):
        st = synth.make_stream("campus", s, T)
        out = generic.run_stream(st, synth.tracker_params("campus"), st.shape.arm_joints)
        for t, (i, _, _) in enumerate(out):
            counts[k, t] = len(i)
            ids[k, t, :len(i)] = torch.from_numpy(i)
        counters[0] += int(counts[k].sum())
        counters[1] += T
    pdist.barrier()
    pdist.reduce_counters(counters)
    g = pdist.gather_results(dict(count=counts, ids=ids), dst=0)
    # the all-gather bench.py times for BASELINE config 5 (every rank receives every rank's blocks)
    ag = {k: torch.empty((w,) + tuple(v.shape), dtype=v.dtype) for k, v in dict(count=counts, ids=ids).items()}
    pdist.all_gather_results(dict(count=counts, ids=ids), ag)
    mx = pdist.max_over_ranks(float(rank + 1))
    q.put((rank, mine, counters.tolist(), mx, None if g is None else {k: v.numpy() for k, v in g.items()},
           {k: v.numpy() for k, v in ag.items()}))


def test_two_rank_shard_reduce_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, mx0, g0, a0), (r1, m1, c1, mx1, g1, a1) = res
    for k in ("count", "ids"):                                  # all-gather: identical on both ranks, equal to the gather
        assert np.array_equal(a0[k], a1[k]) and np.array_equal(a0[k], g0[k])
    assert m0 == [0, 2, 4] and m1 == [1, 3]                   # round-robin, disjoint, complete
    assert c0 == c1 and c0[1] == 5 * 25                        # all-reduced counters agree on every rank
    assert mx0 == mx1 == 2.0
    assert g1 is None and g0["count"].shape == (2, 3, 25) and g0["ids"].shape == (2, 3, 25, 8)
    assert int(g0["count"].sum()) == c0[0]
