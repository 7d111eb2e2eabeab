"""BASELINE.json configs 1 and 2 at FULL size against the CPU oracle, values and every gradient (the golden / seeded
tests of test_gpu_model.py run the same comparison on small meshes and narrow networks).

Protocol as in test_model_matches_oracle_on_seeded_meshes: the fp64 oracle replays the discrete choices (ReLU signs,
max-pool winners) of the CUDA forward, which makes it a smooth function of the inputs; that is the truth.  At these
sizes (ngf 64, 4 levels: widths up to 1024, 19 blocks, ~10^8 ReLU decisions) ANY fp32 evaluation drifts from it by more
than 1e-5 on some tensors -- the reference's own arithmetic included -- so next to the truth the reference-order fp32
oracle is run on the same decisions and the bar per tensor -- outputs, loss and every gradient -- is
    err_cuda <= 1e-5 + err_fp32_oracle
the slack rule of the golden tests (conftest.assert_grads_close): 1e-5 on top of the reference-order fp32 evaluation's
own distance from exact arithmetic (measured: cfg2 outputs 8.3e-6 on both sides, worst gradient tensor 1.1e-5 against
3.4e-6; cfg1 outputs 1.1e-5 on both sides)."""
import pytest
import torch

from conftest import cuda_decisions, rel_err
from oracle import stinet_oracle as O
from test_gpu_model import DECISION_MARGIN, _loss, _oracle_run

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5

CONFIGS = {
    # configs[1]: icosphere subdiv 6 (40,962 vertices), 10 channels, 4 trace-map levels, ngf 64, 9 blocks; two of the
    # eight crops of the batch (per-graph norms, so the batch size does not enter the arithmetic of a graph)
    "cfg2": ("icosphere", dict(subdiv=6), 2,
             dict(input_nc=10, filter_type="edgeconvtransinv", ngf=64, n_blocks=9, n_levels=4)),
    # configs[0]: 4 x 128x128 image-grid graphs, 4 pool levels, EdgeConv, the whole batch
    "cfg1": ("grid", dict(size=128), 4, dict(input_nc=4, filter_type="edgeconv", ngf=64, n_blocks=9, n_levels=4)),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_config_full_size_matches_oracle(name):
    from stinet_b200 import synthetic
    from stinet_b200.models import surfacetextureinpaintingnet as S
    kind, gen_kw, bsz, net_kw = CONFIGS[name]
    torch.manual_seed(49)
    kw = dict(output_nc=3, norm="instance", pooling_type="max", **net_kw)
    net = S.define_G(**kw)
    orc = O.OracleSTINet(**{("norm_type" if k == "norm" else k): v for k, v in kw.items()})
    orc.load_state_dict(net.state_dict())
    batch = synthetic.make_batch(kind, bsz, net_kw["n_levels"], seed=49, **gen_kw)
    net = net.to(DEV)
    gb = batch.to(DEV)
    gb.x = gb.x.clone().requires_grad_(True)
    with cuda_decisions() as cd:
        out = net(gb)
    loss = _loss(out, gb)
    loss.backward()
    gb._stinet_cache.check_status()
    t_out, t_loss, t_grads, dec = _oracle_run(orc, batch, torch.float64, cd.choices)
    r_out, r_loss, r_grads, _ = _oracle_run(orc, batch, torch.float32, cd.choices)
    assert dec.pos == len(cd.choices)
    assert dec.max_relu_margin <= DECISION_MARGIN and dec.max_pool_margin <= DECISION_MARGIN
    e_out, e_ref = rel_err(out, t_out), rel_err(r_out, t_out)
    print(f"{name}: output err vs fp64 truth: cuda {e_out:.2e}, fp32 oracle {e_ref:.2e}; decisions differing from the "
          f"free-running fp64 oracle: relu {dec.n_relu_diff}, pool {dec.n_pool_diff}")
    assert e_out <= TOL + e_ref
    assert rel_err(loss, t_loss) <= TOL + rel_err(r_loss, t_loss)
    g_grads = {k: p.grad for k, p in net.named_parameters()}
    g_grads["__x__"] = gb.x.grad
    scale = max(float(v.abs().max()) for v in t_grads.values())
    worst = (0.0, 0.0, "")
    for k, t in t_grads.items():
        if float(t.abs().max()) < 1e-4 * scale:                      # structurally zero (bias in front of a norm): noise only
            assert float(g_grads[k].abs().max()) < 1e-4 * scale, k
            continue
        e_c, e_r = rel_err(g_grads[k], t), rel_err(r_grads[k], t)
        worst = max(worst, (e_c, e_r, k))
        assert e_c <= TOL + e_r, f"{k}: cuda {e_c:.2e} vs fp32 oracle {e_r:.2e}"
    print(f"{name}: worst gradient tensor {worst[2]}: cuda {worst[0]:.2e}, fp32 oracle {worst[1]:.2e}")
