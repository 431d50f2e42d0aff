"""debug: per-step loss dicts of the harness arm (eager and graph) and the reference-graph arm."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "maskrcnn-benchmark_b200")]
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
which = sys.argv[1] if len(sys.argv) > 1 else "harness"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
if which == "harness":
    from mrb_b200.model import RCNNConfig, build_model
    from mrb_b200.model.backend import B200Backend
    from mrb_b200.optim import ParamArena
    torch.manual_seed(0)
    fixed = os.environ.get("FIXED", "1") == "1"
    cfg = RCNNConfig(mask_rois_per_image=128 if fixed else 0, parallel_heads=os.environ.get("PH", "1") == "1")
    be = B200Backend()
    model = build_model(cfg, backend=be, device=dev).train()
    opt = ParamArena(model.named_parameters(), be, lr=1e-4, momentum=0.9, weight_decay=1e-4)
    be.enable_overlap(os.environ.get("OVERLAP", "1") == "1")
    batches = [bench.synth_batch(2, i, device=dev) for i in range(4)]
    sizes = [(800, 1333)] * 2
    for i in range(steps):
        im, bx, lb = batches[i % 4]
        losses = model(im, sizes, bench.targets_of(bx, lb))
        loss = sum(losses.values())
        loss.backward()
        opt.sync(); opt.step()
        d = {k: round(float(v), 4) for k, v in losses.items()}
        gn = float(opt.param.abs().max())
        print(i, round(float(loss), 4), d, "pmax", round(gn, 3), flush=True)
else:
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
    if which == "refgraph":
        print(fuse_model(model, be))
        opt = ParamArena(model.named_parameters(), be, lr=1e-4, momentum=0.9, weight_decay=1e-4)
        be.enable_overlap(os.environ.get("OVERLAP", "1") == "1")
    else:   # unfused, torch SGD
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-4, momentum=0.9, weight_decay=1e-4)
    data = [bench._ref_inputs(2, i, dev) for i in range(4)]
    for i in range(steps):
        il, tg = data[i % 4]
        losses = model(il, tg)
        loss = sum(losses.values())
        if which != "refgraph":
            opt.zero_grad()
        loss.backward()
        if which == "refgraph":
            opt.sync()
        opt.step()
        print(i, round(float(loss), 4), {k: round(float(v), 4) for k, v in losses.items()}, flush=True)
