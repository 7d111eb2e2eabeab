"""Mint golden vectors from the reference's OWN model code (run in the authoring container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt

/root/reference is imported UNMODIFIED (models/surfacetextureinpaintingnet.py, models/modules/*,
utils/data_utils.py::HierarchicalData); the un-installable third-party modules it needs come from
tests/golden/pyg_shim (see its README).  Inputs are the seeded synthetic graphs of stinet_b200.synthetic.
Each fixture stores: constructor kwargs, state_dict, the batched sample (as collated by the reference's
HierarchicalData rules), the forward output, the trainer loss (inpainting3d_trainer.py:127-137), the gradient of
every parameter and of sample.x, and the argmax of every max-pool (captured by wrapping scatter_max).
The GPU box has no /root/reference: tests only read the .pt files.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "surface-texture-inpainting-net_b200"))
sys.path.insert(0, os.path.join(HERE, "pyg_shim"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
from torch_geometric.data import Batch  # noqa: E402  (shim)
from models import surfacetextureinpaintingnet as ref  # noqa: E402  (REAL reference code)
from utils import data_utils  # noqa: E402  (REAL reference code)
from stinet_b200 import synthetic  # noqa: E402

torch.set_num_threads(4)

CASES = {
    # name: (net kwargs, [per-sample generator specs])
    "edgeconv_grid_b2": (
        dict(input_nc=4, output_nc=3, ngf=8, filter_type="edgeconv", norm="instance", n_blocks=2, n_levels=2,
             pooling_type="max", dilations=[1, 1]),
        [("grid", dict(size=16, n_levels=2, seed=49)), ("grid", dict(size=16, n_levels=2, seed=50))]),
    "edgeconvtransinv_ico_b2": (
        dict(input_nc=10, output_nc=3, ngf=8, filter_type="edgeconvtransinv", norm="instance", n_blocks=3, n_levels=2,
             pooling_type="max", checkpoint_bottleneck=True, dilations=[1, 1, 1]),
        [("icosphere", dict(subdiv=2, n_levels=2, seed=49, mask_radius=2)),
         ("icosphere", dict(subdiv=2, n_levels=2, seed=50, mask_radius=2))]),
    "edgeconvtransinv_ico_dil_b1": (
        dict(input_nc=10, output_nc=3, ngf=8, filter_type="edgeconvtransinv", norm="instance", n_blocks=4, n_levels=2,
             pooling_type="max", checkpoint_bottleneck=True, dilations=[1, 2, 4, 1]),
        [("icosphere", dict(subdiv=3, n_levels=2, seed=51, mask_radius=3, dilations=(2, 4)))]),
    "edgeconv_ragged_b2_meanpool": (
        dict(input_nc=10, output_nc=3, ngf=8, filter_type="edgeconv", norm="instance", n_blocks=1, n_levels=2,
             pooling_type="mean", dilations=[1]),
        [("icosphere", dict(subdiv=3, n_levels=2, seed=52, mask_radius=3)),
         ("icosphere", dict(subdiv=2, n_levels=2, seed=53, mask_radius=2))]),
    "sageconvtransinv_ico_b2": (
        dict(input_nc=10, output_nc=3, ngf=8, filter_type="sageconvtransinv", norm="instance", n_blocks=2, n_levels=1,
             pooling_type="max", dilations=[1, 1]),
        [("icosphere", dict(subdiv=2, n_levels=1, seed=54, mask_radius=2)),
         ("icosphere", dict(subdiv=2, n_levels=1, seed=55, mask_radius=2))]),
    "sageconv_grid_b1": (
        dict(input_nc=4, output_nc=3, ngf=8, filter_type="sageconv", norm="instance", n_blocks=1, n_levels=1,
             pooling_type="max", dilations=[1]),
        [("grid", dict(size=8, n_levels=1, seed=56))]),
    "edgeconv_graphnorm_b2": (
        dict(input_nc=4, output_nc=3, ngf=8, filter_type="edgeconv", norm="graph", n_blocks=1, n_levels=1,
             pooling_type="max", dilations=[1]),
        [("grid", dict(size=8, n_levels=1, seed=57)), ("grid", dict(size=8, n_levels=1, seed=58))]),
    "edgeconv_nonorm_b1": (
        dict(input_nc=4, output_nc=3, ngf=8, filter_type="edgeconv", norm="none", n_blocks=1, n_levels=1,
             pooling_type="max", dilations=[1]),
        [("grid", dict(size=8, n_levels=1, seed=59))]),
}


def to_reference_sample(s):
    """GraphBatch (ours) -> the reference's HierarchicalData, built the way its datasets do
    (scannetcolorgraph_dataloader.py:113-151 / imagegraph_dataloader.py:141-160)."""
    d = data_utils.HierarchicalData(x=s.x, color=s.color, mask=s.mask, edge_index=s.edge_index, name=s.name)
    for k in s.keys:
        if k.startswith("hierarchy_"):
            setattr(d, k, s[k])
    d.num_vertices = s.num_vertices
    return d


def main():
    gens = {"grid": synthetic.grid_sample, "icosphere": synthetic.icosphere_sample}
    for name, (kwargs, specs) in CASES.items():
        torch.manual_seed(49)
        net = ref.define_G(**kwargs)
        net.train()
        samples = [gens[k](**kw) for k, kw in specs]
        batch = Batch.from_data_list([to_reference_sample(s) for s in samples])
        batch.x.requires_grad_(True)

        pool_args = []
        real_scatter_max = ref.scatter_max

        def recording_scatter_max(src, index, dim=0, dim_size=None, **kw):
            out, arg = real_scatter_max(src, index, dim=dim, dim_size=dim_size, **kw)
            if src.is_floating_point():
                pool_args.append(arg.detach().clone())
            return out, arg

        ref.scatter_max = recording_scatter_max
        try:
            out = net(batch)
        finally:
            ref.scatter_max = real_scatter_max
        # trainers/inpainting3d_trainer.py:127-137
        composed = torch.where((batch.mask > 0).expand_as(batch.color), out, batch.color)
        loss = torch.nn.L1Loss(reduction="none")(composed, batch.color)
        loss = loss * torch.pow(0.99, batch.mask.squeeze().float()).unsqueeze(1)
        loss = loss.mean()
        loss.backward()

        fix = {
            "kwargs": kwargs,
            "specs": specs,
            "state_dict": {k: v.detach().clone() for k, v in net.state_dict().items()},
            "sample": {k: (batch[k].detach().clone() if torch.is_tensor(batch[k]) else batch[k])
                       for k in batch.keys if k not in ("ptr", "num_graphs")},
            "out": out.detach().clone(),
            "loss": loss.detach().clone(),
            "grads": {k: p.grad.detach().clone() for k, p in net.named_parameters()},
            "grad_x": batch.x.grad.detach().clone(),
            "pool_args": pool_args,
        }
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(fix, path)
        print(f"{name}: N0={batch.x.shape[0]} out={tuple(out.shape)} loss={loss.item():.6f} "
              f"params={sum(p.numel() for p in net.parameters())} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
