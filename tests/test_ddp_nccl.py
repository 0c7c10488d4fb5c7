"""Multi-GPU parity of the data-parallel training step: N ranks over NCCL, each on its shard of one batch,
against ONE process that runs the same kernels shard by shard (BatchNorm statistics per shard, exactly the per-replica statistics of the
reference's nn.DataParallel, train.py:65-71) and averages -- which equals the gradient of the reference's loss on the gathered batch
(train.py:344-347; model/__init__.py:162-166: every term is a sum over cnt = B_global * cells * A, the class term a mean over the
positives of the WHOLE batch).  Needs >= 2 GPUs (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl.py -m gpu`.
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
    dnn, anchors = _build(rank, load_reference_weights=(rank == 0))
    inference = model.Inference(_config(), dnn, anchors)
    inference = yb_train.ensure_model(inference).train()           # broadcasts rank 0's parameters / buffers
    cfg = _config()
    start = {k: v.detach().clone() for k, v in dnn.state_dict().items()}
    shard = _shard(_batch(), rank)
    opt = torch.optim.SGD(dnn.parameters(), LR)
    res = yb_train.iterate(inference, opt, anchors, cfg, shard)       # kept alive on purpose: iterate() must not hand out the autograd graph
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().float().cpu() for n, p in dnn.named_parameters()}
    after = {k: v.detach().float().cpu() for k, v in dnn.state_dict().items()}
    # the same step from the same start, whole iteration (collectives included) replayed as one CUDA graph
    dnn.load_state_dict(start)
    opt2 = torch.optim.SGD(dnn.parameters(), LR)
    graphed = yb_train.GraphedStep(inference, opt2, anchors, cfg)
    graphed(shard)
    torch.cuda.synchronize()
    after_graph = {k: v.detach().float().cpu() for k, v in dnn.state_dict().items()}
    reducer = ddp.default_reducer(create=False)
    torch.save(dict(grads=grads, after=after, after_graph=after_graph, start={k: v.float().cpu() for k, v in start.items()},
                     loss={k: float(v.item()) for k, v in res['loss'].items()}, buckets=len(dnn.trainer.arena.buckets),
                     bytes=reducer.bytes_reduced, npos=int(res['debug']['pos_count'].sum().item())), os.path.join(out, 'rank%d.pt' % rank))
    graphed.close()
    del graphed
    ddp.shutdown()
    dist.destroy_process_group()


def test_two_rank_step_matches_sharded_single_process():
    if torch.cuda.device_count() < WORLD:
        pytest.skip('needs %d GPUs' % WORLD)
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
    import shutil
    shutil.rmtree(out, ignore_errors=True)
    sd0 = O.make_state_dict(0)
    # rank 0's weights reached rank 1 before the step
    for r in range(WORLD):
        for k, v in res[r]['start'].items():
            if k in sd0:
                assert torch.equal(v, sd0[k].float()), (r, k)
    # every rank ends the step with the same averaged gradients and the same parameters
    for n, g0 in res[0]['grads'].items():
        assert torch.equal(g0, res[1]['grads'][n]), n
    for k in res[0]['after']:
        if 'running' not in k and 'num_batches' not in k:
            assert torch.equal(res[0]['after'][k], res[1]['after'][k]), k
    # ---- the same kernels, one process, shard by shard ----
    import model
    import train as yb_train
    from b200 import ddp
    npos = [res[r]['npos'] for r in range(WORLD)]
    batch = _batch()
    g_sum, running = None, {}
    for r in range(WORLD):
        dnn, anchors = _build(0, True)
        inference = model.Inference(_config(), dnn, anchors).cuda().train()
        # the class term of rank r enters the global mean with weight N_r / N_total (b200.ddp.global_mean_factor)
        cfg = _config(cls_weight=npos[r] / float(sum(npos)))
        opt = torch.optim.SGD(dnn.parameters(), 0.0)
        with ddp.local_only():
            yb_train.iterate(inference, opt, anchors, cfg, _shard(batch, r), reducer=False)
        torch.cuda.synchronize()
        g = {n: p.grad.detach().float().cpu() for n, p in dnn.named_parameters()}
        g_sum = g if g_sum is None else {n: g_sum[n] + g[n] for n in g}
        running[r] = {k: v.float().cpu() for k, v in dnn.state_dict().items() if 'running' in k}
    worst = 0.0
    for n, g in g_sum.items():
        exp = g / WORLD
        got = res[0]['grads'][n]
        e = ((got - exp).norm() / exp.norm().clamp_min(1e-30)).item()
        worst = max(worst, e)
        assert e <= 1e-3, 'averaged gradient %s: rel L2 %.3e' % (n, e)
        p_exp = sd0[n].float() - LR * exp
        e_p = ((res[0]['after'][n] - p_exp).abs().max() / p_exp.abs().max()).item()
        assert e_p <= 1e-5, 'post-step weight %s: %.3e' % (n, e_p)
        e_g = ((res[0]['after_graph'][n] - p_exp).abs().max() / p_exp.abs().max()).item()
        assert e_g <= 1e-5, 'post-step weight through the captured step %s: %.3e' % (n, e_g)
    # BatchNorm statistics stay per rank (the per-replica statistics of DataParallel)
    for r in range(WORLD):
        for k, v in running[r].items():
            assert ((res[r]['after'][k] - v).abs().max() / v.abs().max()).item() <= 1e-4, (r, k)
    assert res[0]['bytes'] >= 4 * sum(g.numel() for g in g_sum.values())
    print('2-rank step: worst averaged-gradient rel L2 %.3e over %d tensors, %d buckets' % (worst, len(g_sum), res[0]['buckets']))
