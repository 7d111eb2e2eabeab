"""Diagnostic: is the (64,64,grid_shuffled) x.grad mismatch a ReLU mask flip at a pre-activation within rounding of 0?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200")); sys.path.insert(0, ROOT)
import torch
from oracle import stinet_oracle as O
from stinet_b200 import synthetic
from stinet_b200.models.modules import edge_conv_filter

s = synthetic.grid_sample(32, 1, seed=5)
ei, n = s.edge_index, s.num_nodes
din, dout = 64, 64
torch.manual_seed(2)
conv = edge_conv_filter.get_gcn_filter(din, dout, module=None, double_input=True)
with torch.no_grad():
    for p in conv.parameters():
        if p.dim() == 1:
            p.normal_(0, 0.5)
x = torch.randn(n, din); go = torch.randn(n, dout)
# fp64 literal pre-activations
xd64 = x.double()
xi, xj = xd64[ei[1]], xd64[ei[0]]
pre = torch.cat([xi, xj - xi], 1) @ conv.nn[0].weight.double().t() + conv.nn[0].bias.double()
print("min |pre| fp64:", pre.abs().min().item(), "count |pre|<1e-6:", int((pre.abs() < 1e-6).sum()), "of", pre.numel())
# fp32 literal
pre32 = torch.cat([x[ei[1]], x[ei[0]] - x[ei[1]]], 1) @ conv.nn[0].weight.t() + conv.nn[0].bias
print("sign flips literal32 vs 64:", int(((pre32 > 0) != (pre > 0)).sum()))
xr = x.clone().requires_grad_(True)
ref = O.edge_conv(xr, ei, conv.nn, "mean", False); ref.backward(go)
conv = conv.cuda()
xg = x.cuda().requires_grad_(True)
out = conv(xg, ei.cuda()); out.backward(go.cuda())
# hoisted pre on GPU
wcat, bcat = conv.hoisted_first_layer()
pq = (xg.detach() @ wcat.t() + bcat)
preg = pq[ei[1].cuda(), :128] + pq[ei[0].cuda(), 128:]
fl = ((preg.cpu() > 0) != (pre > 0))
print("sign flips hoisted-gpu vs fp64:", int(fl.sum()), "at", fl.nonzero().tolist()[:5], "values", pre[fl].tolist()[:5])
d = (xg.grad.cpu() - xr.grad).abs()
rowerr = d.max(1)[0]
print("x.grad max err", d.max().item(), "rows with err>1e-6:", (rowerr > 1e-6).nonzero().flatten().tolist()[:10])
if fl.any():
    e = fl.nonzero()[0, 0].item(); print("flipped edge src,dst:", ei[0, e].item(), ei[1, e].item())
