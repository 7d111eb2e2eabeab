"""Per-tensor error of the one-pass fp16-plane mode ('f16') against the fp64 oracle (decision replay), for the forward
output and every gradient; STINET_F16_BWD_PASSES=3 runs the backward GEMMs in three passes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import cuda_decisions, rel_err
from oracle import stinet_oracle as O
from test_gpu_model import _loss, _oracle_run
from stinet_b200 import synthetic
from stinet_b200.models import surfacetextureinpaintingnet as S
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
for kind, gen_kw, bsz, net_kw in [("icosphere", dict(subdiv=4, mask_radius=4), 3, dict(input_nc=10, filter_type="edgeconvtransinv", ngf=16, n_blocks=2, n_levels=3)),
                                  ("icosphere", dict(subdiv=5, mask_radius=8), 2, dict(input_nc=10, filter_type="edgeconvtransinv", ngf=64, n_blocks=9, n_levels=3))]:
    torch.manual_seed(49)
    kw = dict(output_nc=3, norm="instance", pooling_type="max", **net_kw)
    net = S.define_G(**kw, precision=prec)
    orc = O.OracleSTINet(**{("norm_type" if k == "norm" else k): v for k, v in kw.items()})
    orc.load_state_dict(net.state_dict())
    batch = synthetic.make_batch(kind, bsz, net_kw["n_levels"], seed=49, **gen_kw)
    net = net.to("cuda"); gb = batch.to("cuda")
    with cuda_decisions() as cd:
        out = net(gb)
    _loss(out, gb).backward()
    t_out, t_loss, t_grads, dec = _oracle_run(orc, batch, torch.float64, cd.choices)
    errs = {k: rel_err(p.grad, t_grads[k]) for k, p in net.named_parameters() if float(t_grads[k].abs().max()) > 1e-4 * max(float(v.abs().max()) for v in t_grads.values())}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(prec, "bwd_passes", os.environ.get("STINET_F16_BWD_PASSES", "-"), kind, net_kw["ngf"], "out", f"{rel_err(out, t_out):.2e}", "max grad err", f"{max(errs.values()):.2e}",
          "worst:", [(k, f"{v:.1e}") for k, v in worst], flush=True)
