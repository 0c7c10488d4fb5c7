"""Diagnostic (not a test): one eager train.iterate on the legacy default stream, then train.GraphedStep on the SAME model."""
import configparser
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'yolo2-pytorch_b200')); sys.path.insert(0, ROOT)
import torch  # noqa: E402
import model  # noqa: E402
import model.yolo2  # noqa: E402
import train as yb_train  # noqa: E402
from oracle import yolo2_oracle as O  # noqa: E402

cfg = configparser.ConfigParser()
cfg.read_dict({'batch_norm': {'enable': '1'}, 'model': {'threshold': '0.6'}, 'hparam': {'foreground': '5', 'background': '1', 'center': '1', 'size': '1', 'cls': '1'},
               'train': {'cross_entropy': '1'}})
anchors = O.anchors_yolo_voc()
dnn = model.yolo2.Darknet(model.ConfigChannels(cfg), anchors, 20)
dnn.load_state_dict(O.make_state_dict(0), strict=False)
inference = model.Inference(cfg, dnn, anchors).cuda().train()
x = O.synth_images(6, 160, 160, seed=40)
t = O.synth_targets(6, 160, 160, slots=5, seed=41)
batch = dict(tensor=x, yx_min=t['yx_min'], yx_max=t['yx_max'], cls=t['cls'])
opt = torch.optim.SGD(dnn.parameters(), 1e-3)
if len(sys.argv) < 2 or sys.argv[1] != 'graph_only':
    yb_train.iterate(inference, opt, anchors, cfg, batch)
    torch.cuda.synchronize()
    print('eager ok')
g = yb_train.GraphedStep(inference, opt, anchors, cfg)
g(batch)
torch.cuda.synchronize()
print('graph ok')
