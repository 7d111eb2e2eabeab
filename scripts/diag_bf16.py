"""bf16 mode error budget on the B200: whole-network outputs / gradients vs the fp64 oracle (decision replay),
max-norm and L2-norm relative errors, for a few precision policies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "surface-texture-inpainting-net_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import cuda_decisions
from oracle import stinet_oracle as O
from test_gpu_model import _loss, _oracle_run
from stinet_b200 import synthetic
from stinet_b200.models import surfacetextureinpaintingnet as S

def errs(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), float((a - b).norm() / b.norm().clamp_min(1e-30))

CASES = [("icosphere", dict(subdiv=4, mask_radius=4), 3, dict(input_nc=10, filter_type="edgeconvtransinv", ngf=16, n_blocks=2, n_levels=3)),
         ("grid", dict(size=64), 2, dict(input_nc=4, filter_type="edgeconv", ngf=32, n_blocks=3, n_levels=2)),
         ("icosphere", dict(subdiv=5, mask_radius=6), 2, dict(input_nc=10, filter_type="edgeconvtransinv", ngf=32, n_blocks=9, n_levels=3))]
for kind, gen_kw, bsz, net_kw in CASES:
    for policy in sys.argv[1:] or ["bf16"]:
        torch.manual_seed(49)
        kw = dict(output_nc=3, norm="instance", pooling_type="max", **net_kw)
        net = S.define_G(**kw, precision=policy)
        orc = O.OracleSTINet(**{("norm_type" if k == "norm" else k): v for k, v in kw.items()})
        orc.load_state_dict(net.state_dict())
        batch = synthetic.make_batch(kind, bsz, net_kw["n_levels"], seed=49, **gen_kw)
        net = net.to("cuda"); gb = batch.to("cuda"); gb.x = gb.x.clone().requires_grad_(True)
        with cuda_decisions() as cd:
            out = net(gb)
        loss = _loss(out, gb); loss.backward()
        t_out, t_loss, t_grads, dec = _oracle_run(orc, batch, torch.float64, cd.choices)
        g = {k: p.grad for k, p in net.named_parameters()}; g["__x__"] = gb.x.grad
        scale = max(float(v.abs().max()) for v in t_grads.values())
        ge = {k: errs(g[k], t_grads[k]) for k in t_grads if float(t_grads[k].abs().max()) >= 1e-4 * scale}
        worst_max = max(ge.items(), key=lambda kv: kv[1][0]); worst_l2 = max(ge.items(), key=lambda kv: kv[1][1])
        print(f"{kind} {net_kw['ngf']}/{net_kw['n_blocks']}/{net_kw['n_levels']} {policy}: out max {errs(out, t_out)[0]:.2e} l2 {errs(out, t_out)[1]:.2e} | "
              f"loss {errs(loss, t_loss)[0]:.2e} | grads worst max {worst_max[1][0]:.2e} ({worst_max[0]}) worst l2 {worst_l2[1][1]:.2e} ({worst_l2[0]})", flush=True)
