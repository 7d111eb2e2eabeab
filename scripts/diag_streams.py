"""Which gradients differ in a captured step with the deferred weight-gradient join, and are they stale?  (scripts/gpu_streams.sh)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from stinet_b200 import ops
from stinet_b200.engine import GraphedTrainStep
import test_gpu_streams as T

batches = [T._make(49), T._make(50), T._make(51)]
os.environ["STINET_STRUCT_SIDE_STREAM"] = "0"

def eager():
    ops._WGRAD_SIDE = 0
    net = T._net()
    out = []
    for b in batches:
        gb = b.to("cuda")
        net.zero_grad(set_to_none=True)
        T._loss(net(gb), gb).backward()
        torch.cuda.synchronize()
        out.append({n: p.grad.detach().clone() for n, p in net.named_parameters()})
    return out

def graphed(wg):
    ops._WGRAD_SIDE = wg
    net = T._net()
    step = GraphedTrainStep(net, T._loss, None, warmup=1)
    out = []
    for b in batches:
        step(b.pin_memory())
        torch.cuda.synchronize()
        out.append({n: p.grad.detach().clone() for n, p in net.named_parameters()})
    return out

ref = eager()
for wg in (1, 2):
    got = graphed(wg)
    print("== wgrad mode", wg)
    for i in range(len(batches)):
        for n in ref[i]:
            if not torch.equal(ref[i][n], got[i][n]):
                stale = i > 0 and torch.equal(ref[i - 1][n], got[i][n])
                d = (ref[i][n] - got[i][n]).abs().max().item() / (ref[i][n].abs().max().item() + 1e-30)
                print(f"  step {i} {n:50s} rel diff {d:.3e} stale_by_one={stale} shape={tuple(ref[i][n].shape)}")
