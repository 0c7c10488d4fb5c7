import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200'))
import torch
from b200 import ops
DEV = 'cuda'
b, h, cin, cout, k, pooled = 4, 16, 64, 128, 3, False
gen = torch.Generator().manual_seed(1)
x = (torch.randn(b, cin, h, h, generator=gen) + 0.3).half().float()
wt = (torch.randn(cout, cin, k, k, generator=gen) * (2.0 / (cin * k * k)) ** 0.5).half().float()
gamma = torch.rand(cout, generator=gen) + 0.5
beta = torch.randn(cout, generator=gen) * 0.1
g_out = (torch.randn(b, cout, h, h, generator=gen) * 0.05).half().float()
xr, wr, gr, br = (t.clone().requires_grad_(True) for t in (x, wt, gamma, beta))
z = torch.nn.functional.conv2d(xr, wr, padding=1); z.retain_grad()
y = torch.nn.functional.leaky_relu(torch.nn.functional.batch_norm(z, None, None, gr, br, True, 0.0, 1e-5), 0.1)
(y * g_out).sum().backward()
xd = x.to(DEV).permute(0, 2, 3, 1).contiguous().half()
w16 = ops.pack_weight_f16(wt.to(DEV))
zd = ops.conv_bn_act(xd, w16, torch.ones(cout, device=DEV), torch.zeros(cout, device=DEV), 1.0)
rows = b * h * h
sums = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
mean, invstd = torch.empty(cout, device=DEV), torch.empty(cout, device=DEV)
ops.call('yb_bn_stats', zd, cout, rows, cout, sums)
ops.call('yb_bn_finalize', sums, rows, cout, 1e-5, 0.01, None, None, mean, invstd)
gd, bd = gamma.to(DEV), beta.to(DEV)
go = g_out.to(DEV).permute(0, 2, 3, 1).contiguous().half()
args = (zd, cout, mean, invstd, gd, bd, 0.1, go, cout, 0, None, 0, 0, b, h, h, cout, 0, sums)
ops.call('yb_bn_act_bwd', 0, *args, None, 0, 1)
dz = torch.empty(b, h, h, cout, dtype=torch.float16, device=DEV)
ops.call('yb_bn_act_bwd', 1, *args, dz, cout, 1)
dzr = z.grad.permute(0, 2, 3, 1)
e = (dz.float().cpu() - dzr).abs()
print('dz max err %.3e / max %.3e ; by channel block of 16:' % (e.max(), dzr.abs().max()), [round(float(e[..., i:i + 16].max()), 5) for i in range(0, cout, 16)])
for name, dzz in (('gpu dz', dz), ('ref dz (fp16)', dzr.half().to(DEV).contiguous())):
    dw_krsc = torch.empty(cout, k, k, cin, dtype=torch.float32, device=DEV)
    ops.call('yb_conv_wgrad', xd, dzz, dw_krsc, b, h, h, cin, cout, k, cin, cout)
    ref = wr.grad.permute(0, 2, 3, 1)
    er = (dw_krsc.cpu() - ref).abs()
    print(name, 'dW max err %.3e / max %.3e' % (er.max(), ref.abs().max()))
    print('  by tap:', [round(float(er[:, r, s].max()), 4) for r in range(k) for s in range(k)])
    print('  by co block:', [round(float(er[i:i + 16].max()), 4) for i in range(0, cout, 16)])
    print('  by ci block:', [round(float(er[..., i:i + 16].max()), 4) for i in range(0, cin, 16)])
