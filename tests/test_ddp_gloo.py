"""N>1 path on CPU: two gloo ranks run the harness train step under DistributedDataParallel
(broadcast_buffers=False, as reference tools/train_net.py:49-54) on different synthetic images; after
backward every rank must hold the same gradients, equal to the mean of the two per-rank gradients
computed without DDP (the path shards by images; the only exchange is the gradient all-reduce)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path[:0] = [%(root)r, os.path.join(%(root)r, "maskrcnn-benchmark_b200"), os.path.join(%(root)r, "tests")]
import torch, torch.distributed as dist
from mrb_b200.model import RCNNConfig, GeneralizedRCNN
from oracle.cpu_backend import CpuCheckerBackend
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=world)
cfg = RCNNConfig(stem_out=8, width_per_group=8, res2_out=32, fpn_out=32, mlp_head_dim=64, mask_conv_layers=(32,),
                 roi_batch_size=32, rpn_batch_size=32, pre_nms_top_n_train=100, post_nms_top_n_train=100,
                 fpn_post_nms_top_n_train=100)
torch.manual_seed(0)
model = GeneralizedRCNN(cfg, CpuCheckerBackend()).train()

def batch(seed):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randn(1, 3, 96, 128, generator=g) * 40
    tg = [{"boxes": torch.tensor([[8., 8, 70, 60], [40, 30, 120, 90]]), "labels": torch.tensor([2, 5])}]
    return imgs, [(96, 128)], tg

def grads(m, seed):
    m.zero_grad()
    torch.manual_seed(1234)          # same sampling randomness on every call
    imgs, sizes, tg = batch(seed)
    sum(m(imgs, sizes, tg).values()).backward()
    return [p.grad.clone() for p in m.parameters() if p.requires_grad]

local = [grads(model, 10 + r) for r in range(world)]          # what each rank would get alone
want = [sum(g[i] for g in local) / world for i in range(len(local[0]))]
ddp = torch.nn.parallel.DistributedDataParallel(model, broadcast_buffers=False)
got = grads(ddp, 10 + rank)
for a, b in zip(got, want):
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
# the explicit single-bucket form used by bench.py (graph-capturable): same result as DDP
from mrb_b200.parallel import FlatGradSync
sync = FlatGradSync(model.parameters(), world)
mine = grads(model, 10 + rank)
sync.sync()
for p, a, b in zip([q for q in model.parameters() if q.requires_grad], got, want):
    torch.testing.assert_close(p.grad, b, rtol=1e-4, atol=1e-6)
    assert p.grad.stride() == p.stride()
# ParamArena (the form bench.py uses): gradients live in one flat buffer; the downstream bucket (RPN + ROI heads) is
# reduced early, the backbone bucket and the biases at sync(); together they must equal the plain sum over ranks
from mrb_b200.optim import ParamArena
arena = ParamArena(model.named_parameters(), None, world_size=world)
assert 0 < arena.split < arena.n_w < arena.grad.numel()
torch.manual_seed(1234)                   # (no zero_grad(): p.grad are the arena's persistent, already-zero views)
imgs, sizes, tg = batch(10 + rank)
sum(model(imgs, sizes, tg).values()).backward()    # autograd accumulates into the views
for p, g_local in zip([q for q in model.parameters() if q.requires_grad], local[rank]):
    torch.testing.assert_close(p.grad, g_local, rtol=1e-5, atol=1e-7)
    assert p.grad.data_ptr() >= arena.grad.data_ptr()
assert list(arena.buckets) == ["backbone.body.layer2.", "backbone.body.layer3.", "backbone.body.layer4.", "backbone.fpn.", "heads", "bias"]
# the order in which the boundary nodes of the B200 backend fire during backward: heads, FPN (+ biases), layer4, layer3;
# the first trainable stage (layer2) is left for sync()
for name in ("heads", "backbone.fpn.", "backbone.body.layer4.", "backbone.body.layer3."):
    arena.reduce_bucket(name)
assert set(arena._pending) == {"heads", "backbone.fpn.", "bias", "backbone.body.layer4.", "backbone.body.layer3."}
arena.sync()
assert not arena._pending
for p, b in zip([q for q in model.parameters() if q.requires_grad], want):
    torch.testing.assert_close(p.grad, b * world, rtol=1e-4, atol=1e-6)      # sum; step() folds in the 1/world
flat = torch.cat([g.reshape(-1) for g in got])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert torch.equal(gathered[0], gathered[1])
dist.barrier()
dist.destroy_process_group()
print("RANK", rank, "OK")
'''


def test_two_rank_gloo_gradients_are_averaged(tmp_path, built_lib):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-3000:])
        assert "RANK %d OK" % r in o
