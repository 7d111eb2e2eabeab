"""Reduced-precision modes vs the fp32 mode of the SAME network at BASELINE sizes (forward only): max-norm / L2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "surface-texture-inpainting-net_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from stinet_b200 import synthetic
from stinet_b200.models import surfacetextureinpaintingnet as S

def errs(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())

for name, kind, bsz, L, gen in [("cfg2", "icosphere", 8, 4, dict(subdiv=6)),
                                ("cfg3", "plane", 1, 4, dict(rows=500, cols=500, mask_cover=0.01)),
                                ("cfg5", "plane", 1, 5, dict(rows=1448, cols=1448, mask_cover=0.0001, mask_radius=8))]:
    torch.manual_seed(49)
    net = S.define_G(input_nc=10, output_nc=3, ngf=64, filter_type="edgeconvtransinv", norm="instance", n_blocks=9,
                     n_levels=L, pooling_type="max", gpu_ids=[torch.device("cuda")])
    b = synthetic.make_batch(kind, bsz, L, seed=49, **gen).to("cuda")
    with torch.no_grad():
        ref = net.set_precision("fp32")(b)
        simt = net.set_precision("fp32_simt")(b)
        print(name, "fp32_simt vs fp32(3xTF32): max %.2e l2 %.2e" % errs(simt, ref), flush=True)
        for prec in ("tf32", "bf16"):
            out = net.set_precision(prec)(b)
            print(name, prec, "vs fp32: max %.2e l2 %.2e" % errs(out, ref), flush=True)
