"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (gradient bucketing + all-reduce + averaging,
image sharding).  The same code runs over NCCL on the GPU box (bench.py --mode train under torchrun)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from b200 import ddp
    torch.manual_seed(0)
    shapes = [(7,), (64, 32, 3, 3), (5, 5), (1024,), (3,)]
    grads = {('g%d' % i): torch.full(s, float(rank + 1)) + torch.arange(int(torch.tensor(s).prod())).reshape(s).float() * 0.001 for i, s in enumerate(shapes)}
    red = ddp.GradientAllReducer(bucket_bytes=20000)      # forces several buckets incl. a multi-tensor one
    for name, g in grads.items():
        red.on_grad(name, g)
    red.finish()
    ok = True
    for i, s in enumerate(shapes):
        exp = torch.full(s, (1.0 + world) / 2.0) + torch.arange(int(torch.tensor(s).prod())).reshape(s).float() * 0.001
        ok = ok and torch.allclose(grads['g%d' % i], exp, atol=1e-6)
    spans = [ddp.shard_range(10, r, world) for r in range(world)]
    ok = ok and spans == [(0, 5), (5, 10)] and red.bytes_reduced == sum(4 * int(torch.tensor(s).prod()) for s in shapes)
    # class-term normalisation: per-rank means re-weighted by global_mean_factor and averaged over ranks == the
    # reference's loss on the gathered batch, (sum CE / N_total) / cnt_global with cnt_global = world * cnt_local
    ce_sum = [3.5, 11.25][rank]          # sum of the cross-entropies of this rank's positives
    npos = [4, 9][rank]
    cnt_local = 32.0
    f = ddp.global_mean_factor(torch.tensor(npos))
    mine = torch.tensor(ce_sum / npos / cnt_local) * f
    dist.all_reduce(mine)
    mine = mine / world                  # what gradient averaging does to a per-rank loss term
    reference = (3.5 + 11.25) / (4 + 9) / (cnt_local * world)
    ok = ok and abs(float(mine) - reference) <= 1e-7 and abs(float(f) - npos / 13.0) <= 1e-7
    ok = ok and float(ddp.global_mean_factor(torch.tensor(0), None)) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradient_allreduce_and_sharding_gloo_world2():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'yolo2-pytorch_b200'))
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_range_covers_everything():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'yolo2-pytorch_b200'))
    from b200 import ddp
    for total in (1, 7, 8, 128, 130):
        for world in (1, 2, 3, 8):
            spans = [ddp.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
