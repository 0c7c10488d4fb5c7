#!/usr/bin/env python
"""bench.py -- 416x416 images/sec of the YOLOv2 hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], "C2"): Darknet-19 416x416 batch-32 inference + anchor decode +
class softmax + threshold filter + NMS + per-class expansion, synthetic images, random-init
weights.  One "step" = one batch of 32 images through the whole chain.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # B200 arm (prints ONE JSON line)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU algorithm (oracle port)

N > 1: launched by torch.distributed.run, one rank per GPU; the path shards by image with no
data-path collective (inference replicas), so scaling is "weak" and `value` is the whole-job rate
(N * 32 * K images / max-over-ranks device time).

Numbers in the JSON line:
  value      images/s, inputs resident in HBM (4 rotating input batches = 265 MB > 126 MB L2), CUDA-graph
             replay of the kernel chain, CUDA events on the launching stream, max over ranks.
  e2e        same metric through the serving API with HOST (pinned) fp32 batches: H2D copy of every
             batch and D2H of the detection arrays inside the timed region (double-buffered).
  roofline   tensor-core roofline of the dominant kernel family (the tcgen05 implicit-GEMM conv, 22
             launches/step): algorithmic FLOPs of those launches / their summed CUDA-event durations
             (an instrumented eager pass in the same run), against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline  the oracle (CPU restatement of the reference, torch fp32 on all host cores) on a
             bounded sample of the same workload; rank 0, N=1 only.
"""
import argparse
import configparser
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'yolo2-pytorch_b200')
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GFLOP_PER_IMAGE_416 = 29.360334848          # sum 2*Cin*Cout*k^2*H*W over the 23 convs
GFLOP_LAYER0_416 = 0.299040768              # layers1.0 (direct CUDA-core kernel, not tcgen05)
METRIC = '416x416 images/sec'


def make_config():
    config = configparser.ConfigParser()
    config.read_dict({'batch_norm': {'enable': '1'}, 'model': {'threshold': '0.6'},
                      'detect': {'threshold': '0.3', 'threshold_cls': '0.005', 'fix': '1', 'overlap': '0.45'}})
    return config


ANCHORS_HW = [[1.73145, 1.3221], [4.00944, 3.19275], [8.09892, 5.05587], [4.84053, 9.47112], [10.0071, 11.2364]]


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops=float(d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1400.0))), hbm=float(d.get('hbm_gbs', 6650.0)),
                    source='MEASURED_PEAKS.json (bf16_tflops_sustained)')
    return dict(tflops=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            parts = [p.strip() for p in r.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_chain(O, sd, anchors, x):
    import torch
    with torch.no_grad():
        feature = O.darknet_forward(sd, x)
        pred = O.decode(feature, anchors)
        prob = O.class_prob(pred)
        out = []
        for bi in range(x.size(0)):
            out.append(O.postprocess(pred['iou'][bi].reshape(-1), pred['yx_min'][bi].reshape(-1, 2), pred['yx_max'][bi].reshape(-1, 2),
                                     prob[bi].reshape(-1, prob.size(-1)), True, 0.3, 0.005, 0.45))
    return out


def cpu_measure(batch, steps, warmup, budget_s):
    """Oracle port timed on all host cores.  Returns (images/s, cores, sample description, ms/step)."""
    import torch
    from oracle import yolo2_oracle as O
    ncpu = os.cpu_count() or 1
    sd = O.make_state_dict(0)
    anchors = O.anchors_yolo_voc()
    # calibrate on 4 images: give the CPU path its best thread count (all cores is torch's default,
    # but oversubscription can hurt), then size the per-step sample so the whole run fits the budget
    x2 = O.synth_images(4, 416, 416, seed=0)
    best = None
    for th in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16)}, reverse=True):
        torch.set_num_threads(th)
        cpu_chain(O, sd, anchors, x2)
        t0 = time.perf_counter()
        cpu_chain(O, sd, anchors, x2)
        dt = (time.perf_counter() - t0) / 4
        if best is None or dt < best[0]:
            best = (dt, th)
    per_img, cores = best
    torch.set_num_threads(cores)
    sample = int(max(1, min(batch, budget_s / max(1, steps + warmup) / per_img)))
    x = O.synth_images(sample, 416, 416, seed=0)
    for _ in range(warmup):
        cpu_chain(O, sd, anchors, x)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_chain(O, sd, anchors, x)
    dt = time.perf_counter() - t0
    return sample * steps / dt, cores, '%d steps x %d images (of the %d-image batch), fp32, torch %d threads' % (steps, sample, batch, cores), dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    value, cores, sample, ms = cpu_measure(args.batch, args.steps, args.warmup, budget_s=150.0)
    line = dict(impl='reference', metric=METRIC, value=value, unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='Darknet-19 416x416 batch-%d inference + decode + softmax + filter + NMS (BASELINE configs[1])' % args.batch,
                            global_batch=args.batch, note='CPU: oracle port of the reference algorithm (torch fp32), bounded sample per step'),
                cpu_baseline=dict(value=value, unit='images/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=value, unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_model(device):
    import torch
    import model
    import model.yolo2
    config = make_config()
    anchors = torch.tensor(ANCHORS_HW, dtype=torch.float32)
    torch.manual_seed(0)
    dnn = model.yolo2.Darknet(model.ConfigChannels(config), anchors, 20)
    g = torch.Generator().manual_seed(1)
    for m in dnn.modules():
        if isinstance(m, torch.nn.BatchNorm2d):      # non-trivial BN so the fused scale/shift is exercised
            m.weight.data = torch.rand(m.num_features, generator=g) + 0.5
            m.bias.data = torch.randn(m.num_features, generator=g) * 0.1
            m.running_mean = torch.randn(m.num_features, generator=g) * 0.1
            m.running_var = torch.rand(m.num_features, generator=g) + 0.5
    dnn = dnn.to(device).eval()
    inference = model.Inference(config, dnn, anchors).eval()
    return config, dnn, inference


def profile_layers(pipe, steps):
    """Instrumented eager pass: CUDA events around every tcgen05 conv launch (a long spin kernel in
    front of each step lets the host run ahead so the events bracket back-to-back kernels)."""
    import torch
    from b200 import ops
    records = []
    orig = ops.conv_bn_act

    def timed(x, w, *a, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        y = orig(x, w, *a, **kw)
        e.record()
        b, h, wd, _ = x.shape
        cout, k, _, cin = w.shape
        records.append((s, e, 2.0 * b * h * wd * cin * cout * k * k, (h, cin, cout, k)))
        return y

    ops.conv_bn_act = timed
    other_records = []
    originals = {}

    def wrap_other(name):
        fn = getattr(ops, name)
        originals[name] = fn

        def timed_other(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            y = fn(*a, **kw)
            e.record()
            other_records.append((name, s, e))
            return y
        setattr(ops, name, timed_other)

    for name in ('conv0_bn_leaky_pool', 'conv0_u8_bn_leaky_pool', 'maxpool2x2', 'reorg_f16', 'decode', 'filter_nms'):
        wrap_other(name)
    step_events = []
    try:
        for _ in range(steps):
            torch.cuda._sleep(int(6e7))   # ~30 ms head start: the host enqueues the whole eager step while the GPU spins
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            pipe._forward(pipe.x[0])
            e.record()
            step_events.append((s, e))
        torch.cuda.synchronize()
    finally:
        ops.conv_bn_act = orig
        for name, fn in originals.items():
            setattr(ops, name, fn)
    others = {}
    for name, s, e in other_records:
        others.setdefault(name, []).append(s.elapsed_time(e) * 1e3)
    others = {k: dict(launches_per_step=len(v) // steps, us_per_step=sum(v) / steps) for k, v in others.items()}
    per_step = len(records) // steps
    layers = []
    for i in range(per_step):
        ms = sum(records[j * per_step + i][0].elapsed_time(records[j * per_step + i][1]) for j in range(steps)) / steps
        fl = records[i][2]
        layers.append(dict(shape='%dx%d cin%d cout%d k%d' % ((records[i][3][0],) * 2 + records[i][3][1:]), us=ms * 1e3, tflops=fl / ms / 1e9))
    conv_ms = sum(l['us'] for l in layers) / 1e3
    conv_flops = sum(records[i][2] for i in range(per_step))
    step_ms = sum(s.elapsed_time(e) for s, e in step_events) / steps
    kernels_ms = conv_ms + sum(v['us_per_step'] for v in others.values()) / 1e3
    return dict(layers=layers, others=others, conv_ms=conv_ms, conv_flops=conv_flops, eager_step_ms=step_ms, kernels_ms=kernels_ms, launches=per_step)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from b200 import ops
    from b200.pipeline import DetectPipeline

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: the B200 arm needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    config, dnn, inference = build_model(device)
    B, H, W = args.batch, args.size, args.size
    slots = 4
    pipe = DetectPipeline(inference, config, B, H, W, slots=slots, lanes=args.lanes, use_graph=not args.no_graph).prepare()
    cur = torch.cuda.current_stream()
    g = torch.Generator().manual_seed(100 + rank)
    host = [torch.rand(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    for s in range(slots):
        pipe.x[s].copy_(torch.rand(B, 3, H, W, generator=g))
    torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        pipe.run(i % slots)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(cur)
    pipe.start_after(start)
    for i in range(args.steps):
        pipe.run(i % slots)
    pipe.wait_all(cur)
    end.record(cur)
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    barrier()
    clocks = sampler.stop() if sampler else None
    graph_launches = pipe.launches_per_run * args.steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the serving API: pinned host batches in, detection arrays out ----
    def measure_e2e(pp, host_batches):
        for i in range(max(4, args.warmup)):
            pp.load(i % slots, host_batches[i % 2]); pp.run(i % slots, fetch=True)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record(cur)
        pp.start_after(s2)
        pp.load(0, host_batches[0])
        for i in range(args.steps):
            if i + 1 < args.steps:
                pp.load((i + 1) % slots, host_batches[(i + 1) % 2])
            pp.run(i % slots, fetch=True)
        pp.wait_all(cur)
        e2.record(cur)
        torch.cuda.synchronize()
        t_ms = s2.elapsed_time(e2)
        barrier()
        if world > 1:
            t = torch.tensor([t_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms

    ms_e2e = measure_e2e(pipe, host)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * 3 * H * W * 4
    d2h = pipe.result_bytes()
    # same call with the frames as the reference's data loader holds them before ToTensor (uint8 HWC, detect.py:142-146):
    # the first conv kernel applies the 1/255 and the layout change, so the host->device copy is 4x smaller
    pipe_u8 = DetectPipeline(inference, config, B, H, W, slots=slots, lanes=args.lanes, use_graph=not args.no_graph, uint8_input=True).prepare()
    host_u8 = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    ms_e2e_u8 = measure_e2e(pipe_u8, host_u8)
    e2e_u8 = dict(value=world * B * args.steps / (ms_e2e_u8 / 1e3), unit='images/s', h2d_bytes_per_step=B * 3 * H * W, d2h_bytes_per_step=d2h,
                  ms_per_step=ms_e2e_u8 / args.steps, input='uint8 NHWC frames (ToTensor fused into the first conv kernel)')
    del pipe_u8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family (tcgen05 convs), instrumented pass ----
    peaks = measured_peaks()
    prof = profile_layers(pipe, max(3, min(args.steps, 10)))
    achieved = prof['conv_flops'] / (prof['conv_ms'] / 1e3) / 1e12
    graph_step_ms = ms / args.steps
    roofline = dict(bound='tensor', achieved=achieved, peak=peaks['tflops'], unit='TFLOP/s', frac=achieved / peaks['tflops'],
                    traffic=33.6e6, traffic_source='profiles/r01_ncu_infer_full_v2.md: (dram read 642.5 MB + write 96.8 MB) / 22 tcgen05 conv launches, B=32 (ncu --set full)',
                    kernel='conv_igemm_kernel x21 + conv_c32_kernel x1 (the tcgen05 implicit-GEMM conv family, 22 launches/step)', peak_source=peaks['source'],
                    flops_per_launch=prof['conv_flops'] / prof['launches'], us_per_launch=prof['conv_ms'] * 1e3 / prof['launches'],
                    share_of_step=prof['conv_ms'] / prof['kernels_ms'],
                    whole_step_frac=(B * GFLOP_PER_IMAGE_416 / 1e3) / (graph_step_ms / 1e3) / peaks['tflops'])
    # HBM-bound kernels of the step: algorithmic bytes (SURVEY 8d, fp16 activations inside the backbone) / CUDA-event time
    S = H // 32
    boxes = B * S * S * len(ANCHORS_HW)
    alg_bytes = {
        'conv0_bn_leaky_pool': B * 3 * H * W * 4 + B * (H // 2) * (W // 2) * 32 * 2,
        'maxpool2x2': int(1.25 * 2 * B * ((H // 4) ** 2 * 128 + (H // 8) ** 2 * 256 + (H // 16) ** 2 * 512)),
        'reorg_f16': 2 * B * (H // 16) ** 2 * 64 * 2,
        'decode': B * 125 * S * S * 4 + boxes * (1 + 8 + 20 + 20) * 4,
        'filter_nms': boxes * (1 + 4 + 20) * 4,
    }
    hbm_kernels = {}
    for name, nbytes in alg_bytes.items():
        rec = prof['others'].get(name)
        if rec and rec['us_per_step'] > 0:
            gbs = nbytes / (rec['us_per_step'] * 1e-6) / 1e9
            hbm_kernels[name] = dict(us=rec['us_per_step'], alg_bytes=nbytes, achieved_gbs=gbs, frac_of_hbm_peak=gbs / peaks['hbm'])
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'bench_layers.json'), 'w') as f:
        json.dump(dict(prof, graph_step_ms=graph_step_ms, value=value), f, indent=1)

    cpu = None
    if world == 1 and not args.no_cpu:
        v, cores, sample, _ = cpu_measure(B, 3, 1, budget_s=20.0)
        cpu = dict(value=v, unit='images/s', cores=cores, kind='port', sample=sample)

    line = dict(metric=METRIC, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f16', data='synthetic',
                config=dict(workload='Darknet-19 416x416 batch-%d inference + decode + softmax + filter + NMS (BASELINE configs[1])' % B,
                            global_batch=B * world, per_gpu_batch=B, parallelism='replicas x%d (images shard, no collective)' % world,
                            l2='inputs rotate over %d resident batches (%.0f MB > 126 MB L2); ~0.6 GB of activations streamed per step' % (slots, slots * h2d / 1e6),
                            cuda_graph=not args.no_graph, lanes=pipe.lanes, weights='random-init (kaiming) + random BN statistics'),
                clocks=clocks, e2e=dict(value=e2e_value, unit='images/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=ms_e2e / args.steps,
                         input='fp32 NCHW tensors (the reference forward() signature)'),
                e2e_u8=e2e_u8,
                gpu_launches=graph_launches, roofline=roofline, hbm_kernels=hbm_kernels, cpu_baseline=cpu, wall_ms=wall_ms)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def synth_targets(batch, height, width, slots, generator):
    import torch
    yx_min = torch.zeros(batch, slots, 2)
    yx_max = torch.zeros(batch, slots, 2)
    cls = torch.zeros(batch, slots, dtype=torch.long)
    for b in range(batch):
        n = int(torch.randint(1, slots + 1, (1,), generator=generator))
        hw = torch.rand(n, 2, generator=generator) * torch.tensor([height / 2 - 16.0, width / 2 - 16.0]) + 16.0
        lo, hi = hw / 2, torch.tensor([float(height), float(width)]) - hw / 2
        c = lo + torch.rand(n, 2, generator=generator) * (hi - lo)
        yx_min[b, :n], yx_max[b, :n] = c - hw / 2, c + hw / 2
        cls[b, :n] = torch.randint(0, 20, (n,), generator=generator)
    return yx_min, yx_max, cls


def run_train(args):
    """Secondary measurement (BASELINE configs[2]/[3]): 416x416 training step -- train-mode forward, region loss,
    full backward, NCCL gradient all-reduce when N > 1, Adam step -- images/sec."""
    import torch
    import torch.distributed as dist
    import train as yb_train
    from b200 import ddp, ops
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    config, dnn, inference = build_model(device)
    config.read_dict({'hparam': {'foreground': '5', 'background': '1', 'center': '1', 'size': '1', 'cls': '1'},
                      'train': {'cross_entropy': '1'}})
    dnn.train(); inference.train()
    anchors = torch.tensor(ANCHORS_HW, dtype=torch.float32)
    use_graph = (not args.no_graph) and (world == 1 or args.graph_ddp)
    optimizer = torch.optim.Adam(dnn.parameters(), 1e-5, betas=(0.9, 0.999), eps=1e-8, capturable=use_graph, fused=os.environ.get('YB_ADAM_FUSED', '1') != '0')
    B, H, W = args.batch, args.size, args.size
    g = torch.Generator().manual_seed(200 + rank)
    batches = []
    for _ in range(2):
        yx_min, yx_max, cls = synth_targets(B, H, W, 16, g)
        batches.append(dict(tensor=torch.rand(B, 3, H, W, generator=g).to(device), yx_min=yx_min.to(device), yx_max=yx_max.to(device),
                            cls=cls.to(device)))
    reducer = ddp.GradientAllReducer() if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if use_graph:
        # the whole iteration (fwd, loss, bwd, all-reduce, optimizer) replayed as one CUDA graph; the batch is copied
        # into the graph's static input buffers inside the timed region
        graphed = yb_train.GraphedStep(inference, optimizer, anchors, config, reducer)

        def step(batch):
            return graphed(batch)
    else:
        graphed = None

        def step(batch):
            return yb_train.iterate(inference, optimizer, anchors, config, batch, reducer)
    for i in range(args.warmup):
        step(batches[i % 2])
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = graphed.launches if graphed else ops.launch_count
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        out = step(batches[i % 2])
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    launches = (graphed.launches if graphed else ops.launch_count) - launches0
    barrier()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        value = world * B * args.steps / (ms / 1e3)
        peaks = measured_peaks()
        gflop_train = (3 * GFLOP_PER_IMAGE_416 - GFLOP_LAYER0_416) * (H * W) / (416.0 * 416.0)   # 87.78 GFLOP / image at 416x416
        line = dict(metric='%dx%d training images/sec' % (H, W), value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f16', data='synthetic',
                    config=dict(workload='Darknet-19 %dx%d batch-%d training step: train-mode fwd + region loss + bwd%s + Adam (BASELINE configs[2])'
                                % (H, W, B, ' + NCCL gradient all-reduce' if world > 1 else ''), global_batch=B * world, per_gpu_batch=B,
                                parallelism='dp%d' % world, l2='0.6+ GB of activations per step (> L2)', cuda_graph=bool(use_graph)),
                    clocks=clocks, gpu_launches=launches, loss_total=float(out['loss_total'].item()),
                    roofline=dict(bound='tensor', achieved=value * gflop_train / 1e3, peak=peaks['tflops'], unit='TFLOP/s',
                                  frac=value * gflop_train / 1e3 / peaks['tflops'], traffic=None, kernel='whole training step',
                                  peak_source=peaks['source']),
                    allreduce_bytes_per_step=(reducer.bytes_reduced // max(1, reducer.steps)) if reducer else 0)
        print(json.dumps(line), flush=True)
    if world > 1:
        if use_graph:
            # a captured graph that contains NCCL kernels keeps the communicator busy at teardown (observed: the workers
            # never return from destroy_process_group); everything is measured and printed, so leave without the teardown
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=416)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'], help='train: secondary measurement of the training step')
    ap.add_argument('--lanes', type=int, default=2, help='batches in flight per GPU (one CUDA stream + activation plan each)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--graph-ddp', action='store_true', help='train mode, N > 1: capture the step (incl. the NCCL all-reduce) into a CUDA graph too')
    args = ap.parse_args()
    if args.impl == 'reference':
        args.warmup = max(args.warmup, 1)
        run_reference(args)
    elif args.mode == 'train':
        args.warmup = max(args.warmup, 3)
        if args.batch == 32:
            args.batch = 64
        run_train(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_b200(args)


if __name__ == '__main__':
    main()
