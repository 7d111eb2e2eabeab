"""Compare every intermediate gradient of the whole-network backward between two dense-layer precision modes
(default: fp32 = tensor cores vs fp32_simt = FFMA).  Prints, in backward order, the ops whose incoming gradient differs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from stinet_b200 import ops, synthetic
from stinet_b200.models import surfacetextureinpaintingnet as S
from test_gpu_model import CASES, _loss

kind, gen_kw, bsz, net_kw = CASES[int(os.environ.get("CASE", "0"))]
torch.manual_seed(49)
kw = dict(output_nc=3, norm="instance", pooling_type="max", **net_kw)
net = S.define_G(**kw).to("cuda")
batch = synthetic.make_batch(kind, bsz, net_kw["n_levels"], seed=49, **gen_kw)

records = {}
cur = None
counter = [0]


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **k):
        out = orig(*a, **k)
        t = out[0] if isinstance(out, tuple) else out
        idx = counter[0]
        counter[0] += 1
        tag = f"{idx:03d}:{name}:{tuple(t.shape)}"
        records[cur].setdefault("fwd", {})[tag] = t.detach().clone()
        if t.requires_grad:
            t.register_hook(lambda g, tag=tag: records[cur].setdefault("bwd", {}).__setitem__(tag, g.detach().clone()))
        return out
    setattr(ops, name, f)


for n in ("linear", "edge_message", "norm_act_res", "pool_max", "pool_mean", "unpool", "aggregate"):
    wrap(n)

modes = os.environ.get("MODES", "fp32,fp32_simt").split(",")
for m in modes:
    cur = m
    records[m] = {}
    counter[0] = 0
    net.set_precision(m)
    net.zero_grad(set_to_none=True)
    gb = batch.to("cuda")
    gb.x = gb.x.clone().requires_grad_(True)
    _loss(net(gb), gb).backward()
    records[m]["param"] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}


def rel(a, b):
    d = float(b.abs().max())
    return float((a - b).abs().max()) / max(d, 1e-30)


a, b = records[modes[0]], records[modes[1]]
print("== forward outputs (in order)")
for tag in a["fwd"]:
    e = rel(a["fwd"][tag], b["fwd"][tag])
    if e > 2e-6:
        print(f"  {tag:45s} {e:.2e}")
print("== gradients w.r.t. op outputs (backward order)")
for tag in sorted(a["bwd"], reverse=True):
    e = rel(a["bwd"][tag], b["bwd"][tag])
    ga, gb_ = a["bwd"][tag], b["bwd"][tag]
    nbad = int(((ga - gb_).abs() > 1e-4 * gb_.abs().max()).sum())
    print(f"  {tag:45s} {e:.2e}  elements off by >1e-4 of max: {nbad}")
