"""Execution plan of the Darknet-19 forward pass on the B200 kernels.

The plugin module (`model.yolo2.Darknet`) only holds parameters with the reference's state_dict
names; this engine turns them into kernel operands (fp16 KRSC weights, folded BN scale/shift),
owns one set of fp16 NHWC activation buffers per input shape (sized once, reused every call -- HBM
is 180 GB, the whole B=32 @ 416 plan is < 1 GB) and issues the kernel chain on torch's current
stream.  The 1280-channel concat buffer is written in place by the reorg kernel (channels 0..255)
and by layers2.7 (channels 256..1279), so torch.cat (model/yolo2.py:129) never runs.
"""
import torch

from . import ops


class ConvUnit(object):
    """Operands of one `model.yolo2.Conv2d` unit (conv [+BN] [+leaky]); cached per parameter version."""

    def __init__(self, conv, bn, act):
        self.conv, self.bn, self.act = conv, bn, act
        self._wver = None
        self._bver = None
        self.w16 = self.scale = self.shift = None

    @property
    def cout(self):
        return self.conv.weight.shape[0]

    @property
    def cin(self):
        return self.conv.weight.shape[1]

    @property
    def ksize(self):
        return self.conv.weight.shape[2]

    @property
    def slope(self):
        return 0.1 if self.act else 1.0

    def refresh(self, first_layer=False, force=False):
        """Re-derive the kernel operands when the parameters changed.  `force` skips the version check: training
        steps always re-pack, because fused multi-tensor optimizers (torch.optim.Adam(fused=True)) update parameters
        without advancing torch's version counter."""
        w = self.conv.weight
        wver = (w.data_ptr(), w._version)
        if force or wver != self._wver:
            self.w16 = w.detach().contiguous() if first_layer else ops.pack_weight_f16(w.detach().contiguous(), 0)
            self._wver = wver
        if self.bn is not None:
            ts = (self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
            bver = tuple((t.data_ptr(), t._version) for t in ts)
            if bver != self._bver:
                self.scale, self.shift = ops.bn_fold(*(t.detach().contiguous() for t in ts), eps=self.bn.eps)
                self._bver = bver
        else:
            b = self.conv.bias
            bver = None if b is None else (b.data_ptr(), b._version)
            if bver != self._bver or self.scale is None:
                self.scale = torch.ones(self.cout, dtype=torch.float32, device=w.device)
                self.shift = (b.detach().float().contiguous().clone() if b is not None
                              else torch.zeros(self.cout, dtype=torch.float32, device=w.device))
                self._bver = bver


class DarknetPlan(object):
    """Activation buffers for one (batch, height, width)."""

    def __init__(self, units1, units2, unit_pt, units3, pools1, batch, height, width, device):
        f16 = dict(dtype=torch.float16, device=device)
        self.batch, self.height, self.width = batch, height, width
        h, w = height // 2, width // 2
        self.a0 = torch.empty(batch, h, w, units1[0].cout, **f16)
        self.l1 = []   # (out, pooled or None) per layers1 unit after the first
        for u, pooled in zip(units1[1:], pools1[1:]):
            out = torch.empty(batch, h, w, u.cout, **f16)
            if pooled:
                h, w = h // 2, w // 2
                self.l1.append((out, torch.empty(batch, h, w, u.cout, **f16)))
            else:
                self.l1.append((out, None))
        self.h16, self.w16 = h, w
        self.pt = torch.empty(batch, h, w, unit_pt.cout, **f16)
        self.x1_pool = torch.empty(batch, h // 2, w // 2, units1[-1].cout, **f16)
        h, w = h // 2, w // 2
        self.h32, self.w32 = h, w
        self.cat_ch = unit_pt.cout * 4 + units2[-1].cout
        self.cat = torch.empty(batch, h, w, self.cat_ch, **f16)
        self.l2 = [torch.empty(batch, h, w, u.cout, **f16) for u in units2[:-1]]
        self.l3 = torch.empty(batch, h, w, units3[0].cout, **f16)
        self.feature = torch.empty(batch, units3[1].cout, h, w, dtype=torch.float32, device=device)
        # stream-K scratch (partial sums + flags): per plan, because plans are what run concurrently on different streams
        self.workspace = ops.conv_workspace(device)


class DarknetEngine(object):
    def __init__(self, dnn):
        """`dnn` is a model.yolo2.Darknet (parameter holder).  Units are discovered from its
        nn.Sequential containers so channel-pruned checkpoints (model.ConfigChannels) just work."""
        def unit(m):
            return ConvUnit(m.conv, m.bn if m.has_bn else None, m.has_act)

        self.units1, self.pools1, self._k1 = [], [], []
        mods = list(dnn.layers1)
        for i, m in enumerate(mods):
            if m.is_pool:
                continue
            self.units1.append(unit(m))
            self.pools1.append(i + 1 < len(mods) and mods[i + 1].is_pool)
            self._k1.append('layers1.%d' % i)
        self.units2 = [unit(m) for m in dnn.layers2 if not m.is_pool]
        self._k2 = ['layers2.%d' % i for i, m in enumerate(dnn.layers2) if not m.is_pool]
        self.unit_pt = unit(dnn.passthrough)
        self.units3 = [unit(m) for m in dnn.layers3]
        self.plans = {}
        if not self.pools1[0]:
            raise RuntimeError('Darknet: layers1.0 must be followed by MaxPool2d (fused first-layer kernel)')

    def all_units(self):
        return self.units1 + self.units2 + [self.unit_pt] + self.units3

    def plan(self, batch, height, width, device, plan_id=0):
        key = (batch, height, width, str(device), plan_id)
        p = self.plans.get(key)
        if p is None:
            p = DarknetPlan(self.units1, self.units2, self.unit_pt, self.units3, self.pools1, batch, height, width, device)
            self.plans[key] = p
        return p

    def refresh(self, force=False):
        for i, u in enumerate(self.all_units()):
            u.refresh(first_layer=(i == 0), force=force)

    def invalidate(self):
        """Forget every cached operand (called when the module switches between train() and eval())."""
        for u in self.all_units():
            u._wver = None
            u._bver = None

    def forward(self, x, conv_flags=0, ref=False, collect=None, plan_id=0):
        """x: fp32 NCHW [B,3,H,W] on the GPU -> feature fp32 NCHW [B,A*(5+C),H/32,W/32]
        (a plan-owned buffer, overwritten by the next call with the same shape).
        `collect` (dict) receives references to every unit's fp16 NHWC output (tests)."""
        if not x.is_cuda:
            raise RuntimeError('Darknet (B200): input must be a CUDA tensor; there is no CPU fallback')
        u8 = x.dtype == torch.uint8          # raw RGB frames [B,H,W,3]: the kernel applies ToTensor's 1/255
        if u8:
            b, h, w, c = x.shape
        else:
            b, c, h, w = x.shape
        if c != 3 or h % 32 or w % 32:
            raise ValueError('Darknet expects fp32 [B,3,H,W] or uint8 [B,H,W,3] with H, W multiples of 32, got %s' % (tuple(x.shape),))
        x = x.contiguous() if u8 else x.contiguous().float()
        self.refresh()
        p = self.plan(b, h, w, x.device, plan_id)   # plan_id: independent buffer sets for concurrent streams

        def conv(u, src, dst, **kw):
            return ops.conv_bn_act(src, u.w16, u.scale, u.shift, u.slope, out=dst, flags=conv_flags, ref=ref, workspace=p.workspace, **kw)

        u0 = self.units1[0]
        conv0 = ops.conv0_u8_bn_leaky_pool if u8 else ops.conv0_bn_leaky_pool
        cur = conv0(x, u0.w16, u0.scale, u0.shift, u0.slope, out=p.a0)
        if collect is not None:
            collect['layers1.0(pooled)'] = cur
        x1 = None
        last1 = self.units1[-1]
        for u, (out, pooled), key in zip(self.units1[1:], p.l1, self._k1[1:]):
            # layers1.2 (3x3, Cin = 32) has a kernel whose epilogue applies the MaxPool2d that follows it, so the
            # 208x208x64 activation never goes to HBM; tests that inspect every layer (collect / ref) keep the two steps
            fuse_pool = (pooled is not None and u is not last1 and collect is None and not ref and u.cin == 32 and u.ksize == 3
                         and u.cout <= 64 and (conv_flags & (ops.CONV_NO_SMALLK | ops.CONV_C32_IM2COL)) == 0 and conv_flags < 256)
            if fuse_pool:
                ops.conv_bn_act(cur, u.w16, u.scale, u.shift, u.slope, out=pooled, flags=conv_flags | ops.CONV_POOL2X2)
                cur = pooled
                continue
            conv(u, cur, out)
            if collect is not None:
                collect[key] = out
            x1 = out
            cur = ops.maxpool2x2(out, out=pooled) if pooled is not None else out
        # passthrough branch -> channels [0, 4*Cpt) of the concat buffer
        conv(self.unit_pt, x1, p.pt)
        if collect is not None:
            collect['passthrough'] = p.pt
        ops.reorg_f16(p.pt, p.cat, 0)
        # trunk
        cur = ops.maxpool2x2(x1, out=p.x1_pool)
        for u, out, key in zip(self.units2[:-1], p.l2, self._k2):
            conv(u, cur, out)
            if collect is not None:
                collect[key] = out
            cur = out
        conv(self.units2[-1], cur, p.cat, y_ch_off=self.unit_pt.cout * 4)
        conv(self.units3[0], p.cat, p.l3)
        conv(self.units3[1], p.l3, p.feature, out_mode=ops.OUT_F32_NCHW)
        if collect is not None:
            collect['cat'] = p.cat
            collect['layers3.0'] = p.l3
        return p.feature
