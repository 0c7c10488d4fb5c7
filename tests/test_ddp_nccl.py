"""Multi-GPU parity of the data-parallel training step: N ranks over NCCL, each on its shard of one batch.

What is asserted is what the exchange must guarantee EXACTLY, inside one run: rank 0's parameters reach every rank before the first step
(ranks are built from different seeds); after the step every rank holds the same gradient, and it is the sum of the gradients the ranks
produced locally (snapshotted per bucket on the communication stream right before the all-reduce; 1 / world is folded into the local
values) -- i.e. the average, which for this loss equals the gradient of the reference's loss on the gathered batch (train.py:344-347;
model/__init__.py:162-166: every term is a sum over cnt = B_global * cells * A, the class term a mean over the positives of the WHOLE
batch, reproduced by b200.ddp.global_mean_factor); the SGD step lands on start - lr * average on every rank; the same through the
CUDA-graph-captured step; BatchNorm statistics stay per rank (the per-replica statistics of the reference's nn.DataParallel, train.py:65-71).
A comparison with a separate single-process run is only meaningful loosely: one training step is not reproducible run to run beyond
~1e-2 (head) .. 2e-1 (first layers) at this size (tests/diag_run_to_run.py).
Needs >= 2 GPUs (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl.py -m gpu`.
"""
import configparser
import os
import socket
import time

import pytest
import torch

from oracle import yolo2_oracle as O

pytestmark = pytest.mark.gpu

WORLD, B_RANK, SIZE, LR = 2, 6, 160, 1e-3


def _config(cls_weight=1.0):
    cfg = configparser.ConfigParser()
    cfg.read_dict({'batch_norm': {'enable': '1'}, 'model': {'threshold': '0.6'},
                   'hparam': {'foreground': '5', 'background': '1', 'center': '1', 'size': '1', 'cls': repr(float(cls_weight))},
                   'train': {'cross_entropy': '1'}})
    return cfg


def _batch():
    x = O.synth_images(WORLD * B_RANK, SIZE, SIZE, seed=40)
    t = O.synth_targets(WORLD * B_RANK, SIZE, SIZE, slots=5, seed=41)
    return dict(tensor=x, yx_min=t['yx_min'], yx_max=t['yx_max'], cls=t['cls'])


def _shard(batch, rank):
    from b200 import ddp
    s, e = ddp.shard_range(WORLD * B_RANK, rank, WORLD)
    return {k: v[s:e].clone() for k, v in batch.items()}


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(rank_seed, load_reference_weights):
    import model
    import model.yolo2
    anchors = O.anchors_yolo_voc()
    torch.manual_seed(1000 + rank_seed)                 # ranks start from DIFFERENT random weights ...
    dnn = model.yolo2.Darknet(model.ConfigChannels(_config()), anchors, 20)
    if load_reference_weights:
        dnn.load_state_dict(O.make_state_dict(0), strict=False)     # ... only rank 0 holds the intended ones
    return dnn, anchors


def _worker(rank, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
    import model
    import train as yb_train
    from b200 import ddp

    class SnapshotReducer(ddp.GradientAllReducer):
        """Keeps a copy of every bucket as this rank produced it (taken on the communication stream right before the all-reduce)."""
        snaps = {}

        def _launch(self, bi, streams):
            for st in [s for s in streams if s is not None]:
                ev = torch.cuda.Event()
                ev.record(st)
                self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                self.snaps[bi] = self.arena.bucket_tensor(bi).clone()
            return ddp.GradientAllReducer._launch(self, bi, streams)

    reducer = SnapshotReducer()
    ddp.set_default_reducer(reducer)
    dnn, anchors = _build(rank, load_reference_weights=(rank == 0))
    inference = model.Inference(_config(), dnn, anchors)
    inference = yb_train.ensure_model(inference).train()           # broadcasts rank 0's parameters / buffers
    cfg = _config()
    sd0 = O.make_state_dict(0)
    start = {k: v.detach().clone() for k, v in dnn.state_dict().items()}
    start_ok = all(torch.equal(start[k].float().cpu(), sd0[k].float()) for k in sd0)
    shard = _shard(_batch(), rank)
    opt = torch.optim.SGD(dnn.parameters(), LR)
    res = yb_train.iterate(inference, opt, anchors, cfg, shard)       # kept alive on purpose: iterate() must not hand out the autograd graph
    torch.cuda.synchronize()
    arena = dnn.trainer.arena
    local_flat = torch.cat([reducer.snaps[bi] for bi in range(len(arena.buckets))])
    assert local_flat.numel() == arena.flat.numel()
    grads, local = {}, {}
    for n, p in dnn.named_parameters():
        assert p.grad.data_ptr() == arena.views[n].data_ptr(), n          # .grad IS the arena slot
        off = (arena.views[n].data_ptr() - arena.flat.data_ptr()) // 4
        grads[n] = p.grad.detach().float().cpu()
        local[n] = local_flat[off:off + p.numel()].view_as(p).float().cpu()
    after = {k: v.detach().float().cpu() for k, v in dnn.state_dict().items()}
    # the same step from the same start, whole iteration (collectives included) replayed as one CUDA graph
    dnn.load_state_dict(start)
    opt2 = torch.optim.SGD(dnn.parameters(), LR)
    graphed = yb_train.GraphedStep(inference, opt2, anchors, cfg)
    graphed(shard)
    graphed(shard)
    torch.cuda.synchronize()
    after_graph = {k: v.detach().float().cpu() for k, v in dnn.state_dict().items()}
    torch.save(dict(grads=grads, local=local, after=after, after_graph=after_graph, start_ok=start_ok,
                    loss={k: float(v.item()) for k, v in res['loss'].items()}, buckets=len(arena.buckets),
                    bytes=reducer.bytes_reduced, npos=int(res['debug']['pos_count'].sum().item())), os.path.join(out, 'rank%d.pt' % rank))
    graphed.close()
    del graphed
    ddp.shutdown()
    dist.destroy_process_group()


def test_two_rank_step_exchanges_exact_average():
    if torch.cuda.device_count() < WORLD:
        pytest.skip('needs %d GPUs' % WORLD)
    import shutil
    import tempfile
    import torch.multiprocessing as mp
    out = tempfile.mkdtemp(prefix='yb_ddp_')
    ctx = mp.start_processes(_worker, args=(_free_port(), out), nprocs=WORLD, join=False, start_method='spawn')
    deadline = time.time() + 300
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail('data-parallel workers did not finish within 300 s (hang in the exchange or its teardown)')
    res = {r: torch.load(os.path.join(out, 'rank%d.pt' % r)) for r in range(WORLD)}
    shutil.rmtree(out, ignore_errors=True)
    sd0 = O.make_state_dict(0)
    # rank 0's weights reached rank 1 before the step (rank 1 was built from a different seed)
    assert res[0]['start_ok'] and res[1]['start_ok']
    worst = 0.0
    for n, g0 in res[0]['grads'].items():
        # every rank holds the same gradient after the exchange ...
        assert torch.equal(g0, res[1]['grads'][n]), n
        # ... and it is the sum of what the two ranks produced (1 / world is already folded into the local values): exact for 2 ranks
        exp = res[0]['local'][n] + res[1]['local'][n]
        assert not torch.equal(res[0]['local'][n], res[1]['local'][n]) or n.endswith('num_batches_tracked'), n     # the shards differ
        e = ((g0 - exp).abs().max() / exp.abs().max().clamp_min(1e-30)).item()
        worst = max(worst, e)
        assert e <= 1e-6, 'all-reduced gradient %s: %.3e' % (n, e)
        # SGD step on the averaged gradient: identical parameters on both ranks, equal to start - lr * grad
        p_exp = sd0[n].float() - LR * g0
        for r in range(WORLD):
            assert ((res[r]['after'][n] - p_exp).abs().max() / p_exp.abs().max()).item() <= 1e-6, (r, n)
    for k in res[0]['after_graph']:
        if 'running' in k or 'num_batches' in k:
            continue
        # the captured step (all-reduces inside the CUDA graph): both ranks step identically, and they did step (two replays)
        assert torch.equal(res[0]['after_graph'][k], res[1]['after_graph'][k]), k
        assert not torch.equal(res[0]['after_graph'][k], sd0[k].float()), k
    # BatchNorm statistics stay per rank (the per-replica statistics of DataParallel): the shards differ, so must they
    assert any(not torch.equal(res[0]['after'][k], res[1]['after'][k]) for k in res[0]['after'] if 'running_mean' in k)
    assert res[0]['bytes'] >= 4 * sum(g.numel() for g in res[0]['grads'].values())
    # against ONE process stepping on shard 0 alone (same kernels): the class term enters with weight N_0 / N_total.  One training step is
    # not reproducible run to run beyond ~1e-2 at the head and ~2e-1 at the first layers (fp32 atomics order amplified by 22 train-mode
    # BatchNorm layers, tests/diag_run_to_run.py), so only the head's bias gradient is compared, loosely.
    import model
    import train as yb_train
    from b200 import ddp
    npos = [res[r]['npos'] for r in range(WORLD)]
    dnn, anchors = _build(0, True)
    inference = model.Inference(_config(), dnn, anchors).cuda().train()
    with ddp.local_only():
        yb_train.iterate(inference, torch.optim.SGD(dnn.parameters(), 0.0), anchors, _config(cls_weight=npos[0] / float(sum(npos))), _shard(_batch(), 0),
                         reducer=False)
    torch.cuda.synchronize()
    single = dict(dnn.named_parameters())['layers3.1.conv.bias'].grad.float().cpu() / WORLD
    ddp_local = res[0]['local']['layers3.1.conv.bias']
    e = ((single - ddp_local).norm() / ddp_local.norm()).item()
    assert e <= 5e-2, 'rank 0 local head-bias gradient vs single process on shard 0: rel L2 %.3e' % e
    print('2-rank step: all-reduce exact to %.1e over %d tensors, %d buckets; head-bias local gradient vs single process %.2e' % (worst, len(res[0]['grads']), res[0]['buckets'], e))
