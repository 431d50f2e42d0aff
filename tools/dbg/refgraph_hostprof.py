"""debug: where the HOST time of the reference-graph arm goes (cProfile over 5 eager steps, top cumulative entries)."""
import cProfile, io, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")]
import torch
import bench
dev = torch.device("cuda", 0)
from mrb_b200 import refenv, engine
refenv.activate()
from maskrcnn_benchmark.config import cfg as _cfg
from maskrcnn_benchmark.modeling.detector import build_detection_model
from mrb_b200.fuse import fuse_model
from mrb_b200.model.backend import B200Backend
from mrb_b200.optim import ParamArena
cfg = _cfg.clone(); cfg.merge_from_file(refenv.config_path("e2e_mask_rcnn_R_50_FPN_1x.yaml")); cfg.merge_from_list(["MODEL.DEVICE", "cuda"]); cfg.freeze()
torch.manual_seed(0)
model = build_detection_model(cfg).to(dev).train()
be = B200Backend(); engine.set_default_backend(be)
fuse_model(model, be)
opt = ParamArena(model.named_parameters(), be, lr=1e-4, momentum=0.9, weight_decay=1e-4)
be.enable_overlap(True)
data = [bench._ref_inputs(2, i, dev) for i in range(4)]
if os.environ.get("MRB_REFGRAPH_SEGMENTS", "1") != "0":
    from mrb_b200.graphed import graph_module
    graph_module(model.backbone, (data[0][0].tensors,), backend=be, arena=opt)
    graph_module(model.rpn.head, (list(model.backbone(data[0][0].tensors)),), backend=be, arena=opt, share_inputs=True)
def step(i):
    il, tg = data[i % 4]
    losses = model(il, tg)
    loss = sum(losses.values())
    loss.backward(); opt.sync(); opt.step()
    return loss
for i in range(4):
    step(i)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(5):
    step(i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(90)
print(s.getvalue()[:20000])
