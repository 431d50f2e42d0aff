"""debug: find the first non-finite tensor of the harness train step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")]
import torch
import bench
dev = torch.device("cuda", 0)
from mrb_b200.model import RCNNConfig, build_model
from mrb_b200.model.backend import B200Backend
from mrb_b200.optim import ParamArena
torch.manual_seed(0)
cfg = RCNNConfig(mask_rois_per_image=128, parallel_heads=False)
be = B200Backend()
model = build_model(cfg, backend=be, device=dev).train()
opt = ParamArena(model.named_parameters(), be, lr=1e-4, momentum=0.9, weight_decay=1e-4)
be.enable_overlap(os.environ.get("OVERLAP", "0") == "1")
batches = [bench.synth_batch(2, i, device=dev) for i in range(4)]
sizes = [(800, 1333)] * 2
names = {id(p): n for n, p in model.named_parameters()}
def bad(t):
    return not bool(torch.isfinite(t.float()).all())
from mrb_b200.model import roi_heads, box_ops
orig_loss = roi_heads.BoxHead.loss
def loss_dbg(self, cls, reg, labels, reg_t):
    print("   box loss inputs: cls bad", bad(cls), "reg bad", bad(reg), "reg_t bad", bad(reg_t), "labels range", int(labels.min()), int(labels.max()),
          "n_pos", int((labels > 0).sum()), "reg absmax", float(reg.float().abs().max()), "reg_t absmax", float(reg_t.abs().max()))
    if bad(reg_t):
        idx = (~torch.isfinite(reg_t.reshape(-1, 4)).all(1)).nonzero().squeeze(1)[:5]
        print("   bad reg_t rows", idx.tolist(), reg_t.reshape(-1, 4)[idx].tolist(), "labels", labels.reshape(-1)[idx].tolist())
    return orig_loss(self, cls, reg, labels, reg_t)
roi_heads.BoxHead.loss = loss_dbg
orig_sub = roi_heads.BoxHead.subsample
def sub_dbg(self, proposals, targets, generator=None, be=None):
    boxes, scores, valid = proposals
    print("   proposals: boxes bad", bad(boxes), "absmax", float(boxes.abs().max()), "valid", int(valid.sum()))
    out = orig_sub(self, proposals, targets, generator, be)
    return out
roi_heads.BoxHead.subsample = sub_dbg
for i in range(4):
    im, bx, lb = batches[i % 4]
    losses = model(im, sizes, bench.targets_of(bx, lb))
    loss = sum(losses.values())
    print(i, {k: round(float(v.detach()), 4) for k, v in losses.items()}, flush=True)
    loss.backward()
    g = opt.grad
    if bad(g):
        for p, gv in opt.sinks.values():
            if bad(gv):
                print("   non-finite grad:", names[id(p)], tuple(p.shape)); break
    opt.sync(); opt.step()
    if bad(opt.param):
        for p, gv in opt.sinks.values():
            if bad(p):
                print("   non-finite param after step:", names[id(p)]); break
