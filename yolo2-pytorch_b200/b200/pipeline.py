"""Batched detection pipeline: backbone -> decode (+softmax) -> threshold filter -> NMS -> per-class
expansion, i.e. what the reference's `detect.Detect.__call__` does per frame (detect.py:141-153)
for a whole batch, as one CUDA-graph replay.

Serving layout: `slots` device-resident input batches (fp32 NCHW, the tensor the reference's
callers produce -- or raw uint8 NHWC frames).  Slot s runs on lane `s % lanes`; every lane has its
own CUDA stream and its own activation plan, so with lanes = 2 two batches are in flight at once and
the second batch's kernels fill the SMs the first leaves idle (wave tails of the persistent conv
kernels, the 32-CTA NMS kernel, launch gaps).  `load(slot, host_tensor)` enqueues the host->device
copy on a copy stream, `run(slot)` replays the captured kernel chain on the slot's lane, `fetch(slot)`
copies the (small) detection arrays back to pinned host memory on the same lane.
"""
import torch

from . import ops


class DetectPipeline(object):
    RESULT_KEYS = ('n_filtered', 'n_keep', 'keep_box', 'n_det', 'det_keep', 'det_cls', 'det_score', 'best_cls', 'iou', 'yx_min', 'yx_max')

    def __init__(self, inference, config, batch, height, width, slots=2, lanes=1, use_graph=True, limit=200, device=None, uint8_input=False):
        self.inference = inference
        self.dnn = inference.dnn
        self.engine = self.dnn.engine
        self.config = config
        self.batch, self.height, self.width = batch, height, width
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.limit = limit
        self.fix = config.getboolean('detect', 'fix')
        self.mode = ops.FILTER_FIX if self.fix else ops.FILTER_THRESHOLD
        self.threshold = 0.0 if self.fix else config.getfloat('detect', 'threshold')
        self.threshold_cls = config.getfloat('detect', 'threshold_cls') if self.fix else 0.0
        self.overlap = config.getfloat('detect', 'overlap')
        self.anchors = inference.anchors.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if uint8_input:
            self.x = [torch.zeros(batch, height, width, 3, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        else:
            self.x = [torch.zeros(batch, 3, height, width, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self.lanes = max(1, min(lanes, slots))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
        self.out = [None] * slots
        self.graphs = [None] * slots
        self.host = [None] * slots
        self.use_graph = use_graph
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.loaded = [None] * slots   # event: H2D of this slot finished
        self.done = [None] * slots     # event: kernels (and result copies) of this slot finished
        self.launches_per_run = None

    def lane_of(self, slot):
        return slot % self.lanes

    # ---- the kernel chain ----------------------------------------------------------------------
    def _forward(self, x, plan_id=0):
        feature = self.engine.forward(x, plan_id=plan_id)
        a = self.anchors.size(0)
        per = feature.size(1) // a
        num_cls = per - 5 if per > 5 else 1
        dec = ops.decode(feature, self.anchors, num_cls, with_prob=True)
        b = feature.size(0)
        n = dec['iou'][0].numel()
        res = ops.filter_nms(dec['iou'].view(b, n), dec['yx_min'].view(b, n, 2), dec['yx_max'].view(b, n, 2),
                             dec['prob'].view(b, n, -1), self.mode, self.threshold, self.threshold_cls, self.overlap, self.limit,
                             expand=True, details=True)
        out = dict(feature=feature)
        out.update(dec)
        out.update(res)
        return out

    def prepare(self):
        """Warm up (weight packing, attribute setting, lazy allocations) and capture one graph per slot."""
        cur = torch.cuda.current_stream(self.device)
        for lane, st in enumerate(self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(2):
                    before = ops.launch_count
                    self.out[lane] = self._forward(self.x[lane], plan_id=lane)
                    self.launches_per_run = ops.launch_count - before
            cur.wait_stream(st)
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            for s in range(len(self.x)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.streams[self.lane_of(s)]):
                    self.out[s] = self._forward(self.x[s], plan_id=self.lane_of(s))
                self.graphs[s] = g
            torch.cuda.synchronize(self.device)
        return self

    # ---- serving calls ----------------------------------------------------------------------------
    def load(self, slot, host_tensor):
        """Enqueue the host->device copy of one batch (pinned) on the copy stream."""
        if self.done[slot] is not None:
            self.copy_stream.wait_event(self.done[slot])
        with torch.cuda.stream(self.copy_stream):
            self.x[slot].copy_(host_tensor, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.loaded[slot] = ev

    def run(self, slot, fetch=False):
        """Run the chain on slot `slot` on its lane's stream; returns the dict of device result tensors
        (valid once `self.done[slot]` has completed, e.g. after `sync()`).  fetch=True also enqueues the
        device->host copy of the detection arrays (see `fetch`)."""
        st = self.streams[self.lane_of(slot)]
        if self.loaded[slot] is not None:
            st.wait_event(self.loaded[slot])
            self.loaded[slot] = None
        with torch.cuda.stream(st):
            if self.graphs[slot] is not None:
                self.graphs[slot].replay()
            else:
                self.out[slot] = self._forward(self.x[slot], plan_id=self.lane_of(slot))
            if fetch:
                self._fetch(slot)
            ev = torch.cuda.Event()
            ev.record(st)
        self.done[slot] = ev
        return self.out[slot]

    def _fetch(self, slot):
        out = self.out[slot]
        if self.host[slot] is None:
            self.host[slot] = {k: torch.empty(out[k].shape, dtype=out[k].dtype, pin_memory=True) for k in self.RESULT_KEYS}
        for k in self.RESULT_KEYS:
            self.host[slot][k].copy_(out[k], non_blocking=True)
        return self.host[slot]

    def fetch(self, slot):
        """Device->host copy of the detection arrays of `slot` into pinned buffers, on the slot's lane."""
        st = self.streams[self.lane_of(slot)]
        with torch.cuda.stream(st):
            host = self._fetch(slot)
            ev = torch.cuda.Event()
            ev.record(st)
        self.done[slot] = ev
        return host

    def wait_all(self, stream=None):
        """Make `stream` (default: current) wait for everything enqueued on the lanes."""
        stream = torch.cuda.current_stream(self.device) if stream is None else stream
        for st in self.streams:
            stream.wait_stream(st)

    def start_after(self, event):
        """Make all lanes and the copy stream wait for `event` (timing fences)."""
        for st in self.streams:
            st.wait_event(event)
        self.copy_stream.wait_event(event)

    def result_bytes(self):
        out = self.out[0]
        return sum(out[k].numel() * out[k].element_size() for k in self.RESULT_KEYS)

    def detections(self, host, bi):
        """Per-image view of fetched results, in the form `detect.postprocess` returns
        (iou[k], yx_min[m,2], yx_max[m,2], cls[m], score[m]) or None."""
        nk = int(host['n_keep'][bi])
        if nk == 0:
            return None
        kbox = host['keep_box'][bi, :nk].long()
        iou = host['iou'][bi].reshape(-1)[kbox]
        yx_min, yx_max = host['yx_min'][bi].reshape(-1, 2), host['yx_max'][bi].reshape(-1, 2)
        if not self.fix:
            return iou, yx_min[kbox], yx_max[kbox], host['best_cls'][bi][kbox].long(), iou
        nd = int(host['n_det'][bi])
        dbox = kbox[host['det_keep'][bi, :nd].long()]
        return iou, yx_min[dbox], yx_max[dbox], host['det_cls'][bi, :nd].long(), host['det_score'][bi, :nd]
