"""Per-tensor gradient error of the whole network against the fp64 CPU oracle, for the dense-layer precision modes
(fp32 = 3xTF32 tensor cores, fp32_simt = FFMA) next to the fp32 CPU oracle's own error.  Run on the GPU box."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import stinet_oracle as O
from stinet_b200 import synthetic
from stinet_b200.models import surfacetextureinpaintingnet as S
from test_gpu_model import CASES, _loss, _oracle_run

for kind, gen_kw, bsz, net_kw in CASES[:2]:
    torch.manual_seed(49)
    kw = dict(output_nc=3, norm="instance", pooling_type="max", **net_kw)
    net = S.define_G(**kw)
    orc = O.OracleSTINet(**{("norm_type" if k == "norm" else k): v for k, v in kw.items()})
    orc.load_state_dict(net.state_dict())
    batch = synthetic.make_batch(kind, bsz, net_kw["n_levels"], seed=49, **gen_kw)
    t_out, t_loss, t_grads = _oracle_run(orc, batch, torch.float64)
    c_out, c_loss, c_grads = _oracle_run(orc, batch, torch.float32)
    net = net.to("cuda")
    res = {}
    for prec in ("fp32", "fp32_simt"):
        net.set_precision(prec)
        net.zero_grad(set_to_none=True)
        gb = batch.to("cuda")
        gb.x = gb.x.clone().requires_grad_(True)
        out = net(gb)
        _loss(out, gb).backward()
        g = {k: p.grad.detach().cpu().double() for k, p in net.named_parameters()}
        g["__x__"] = gb.x.grad.detach().cpu().double()
        res[prec] = (float((out.detach().cpu().double() - t_out).abs().max() / t_out.abs().max()), g)
    print(kind, "out err: tc %.2e simt %.2e cpu32 %.2e" % (res["fp32"][0], res["fp32_simt"][0],
          float((c_out.double() - t_out).abs().max() / t_out.abs().max())))
    for k, t in t_grads.items():
        d = float(t.abs().max())
        if d == 0:
            continue
        e = [float((res[p][1][k] - t).abs().max()) / d for p in ("fp32", "fp32_simt")]
        ec = float((c_grads[k].double() - t).abs().max()) / d
        print(f"  {k:50s} {tuple(t.shape)!s:16s} tc {e[0]:.2e} simt {e[1]:.2e} cpu32 {ec:.2e}")
