"""Mint golden vectors for SingleConvMeshNet from the reference's OWN code (authoring container only).

    python tests/golden/make_golden_singleconv.py        # rewrites tests/golden/singleconv/*.pt

/root/reference/models/singleconvmeshnet.py and models/modules/edge_conv_filter.py (with_norm=True: BatchNorm1d over
EDGES inside the message MLP, :34-44) are imported unmodified on top of tests/golden/pyg_shim.  The network is run in
train() mode (batch statistics, running statistics updated) on seeded synthetic meshes; stored: constructor kwargs,
state_dict before the step, the collated sample, output, a scalar loss (mean squared output, the segmentation trainer
is out of scope), every parameter gradient, grad of x, and the BatchNorm buffers after the step.
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
from models import singleconvmeshnet as ref  # noqa: E402  (REAL reference code)
from utils import data_utils  # noqa: E402  (REAL reference code)
from stinet_b200 import synthetic  # noqa: E402

torch.set_num_threads(4)

CASES = {
    "singleconv_ico_mean_b2": (
        dict(feature_number=10, num_propagation_steps=1, filter_sizes=[8, 16, 24], num_classes=5,
             pooling_method="mean", aggr="mean"),
        [("icosphere", dict(subdiv=2, n_levels=2, seed=61, mask_radius=2)),
         ("icosphere", dict(subdiv=2, n_levels=2, seed=62, mask_radius=2))]),
    # two propagation steps: the reference's in-place residual (:107) makes its own backward raise, so forward only
    "singleconv_ico_steps2_fwd": (
        dict(feature_number=10, num_propagation_steps=2, filter_sizes=[8, 12], num_classes=4,
             pooling_method="mean", aggr="mean"),
        [("icosphere", dict(subdiv=2, n_levels=1, seed=64, mask_radius=2))]),
    "singleconv_ico_max_b1": (
        dict(feature_number=10, num_propagation_steps=1, filter_sizes=[12, 20], num_classes=3,
             pooling_method="max", aggr="mean"),
        [("icosphere", dict(subdiv=3, n_levels=1, seed=63, mask_radius=3))]),
}


def to_reference_sample(s):
    d = data_utils.HierarchicalData(x=s.x, color=s.color, mask=s.mask, edge_index=s.edge_index, name=s.name)
    for k in s.keys:
        if k.startswith("hierarchy_"):
            setattr(d, k, s[k])
    d.num_vertices = s.num_vertices
    return d


def main():
    os.makedirs(os.path.join(HERE, "singleconv"), exist_ok=True)
    gens = {"grid": synthetic.grid_sample, "icosphere": synthetic.icosphere_sample}
    for name, (kwargs, specs) in CASES.items():
        torch.manual_seed(49)
        net = ref.SingleConvMeshNet(**kwargs)
        net.train()
        state = {k: v.detach().clone() for k, v in net.state_dict().items()}
        samples = [gens[k](**kw) for k, kw in specs]
        batch = Batch.from_data_list([to_reference_sample(s) for s in samples])
        batch.x.requires_grad_(True)
        out = net(batch)
        loss = out.square().mean()
        forward_only = kwargs["num_propagation_steps"] > 1
        if not forward_only:
            loss.backward()
        fix = {
            "kwargs": kwargs, "specs": specs, "state_dict": state,
            "sample": {k: (batch[k].detach().clone() if torch.is_tensor(batch[k]) else batch[k])
                       for k in batch.keys if k not in ("ptr", "num_graphs")},
            "out": out.detach().clone(), "loss": loss.detach().clone(),
            "forward_only": forward_only,
            "grads": None if forward_only else {k: p.grad.detach().clone() for k, p in net.named_parameters()},
            "grad_x": None if forward_only else batch.x.grad.detach().clone(),
            "buffers_after": {k: v.detach().clone() for k, v in net.named_buffers()},
        }
        path = os.path.join(HERE, "singleconv", f"{name}.pt")
        torch.save(fix, path)
        print(f"{name}: N0={batch.x.shape[0]} out={tuple(out.shape)} loss={loss.item():.6f} "
              f"params={sum(p.numel() for p in net.parameters())} keys={len(state)} -> {os.path.getsize(path) / 1024:.0f} KiB")
        print("  first keys:", list(state)[:6])


if __name__ == "__main__":
    main()
