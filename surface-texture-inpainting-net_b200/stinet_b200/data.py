"""Host-side sample / batch containers for the STINet hot path.

`GraphBatch` is a duck-typed stand-in for a PyG `Batch` of the reference's
`HierarchicalData` (reference utils/data_utils.py:11-42): attribute access (`sample.x`,
`sample.edge_index`, `sample.batch`, `sample.num_vertices`) and string-key access
(`sample["hierarchy_trace_index_2"]`) are the only two things the model touches
(reference models/surfacetextureinpaintingnet.py:404-455).  A real PyG Batch works equally
well as model input; this class exists so the path never imports torch_geometric.

`collate` restates the batching offsets of `HierarchicalData.__inc__/__cat_dim__`
(reference utils/data_utils.py:23-42) on top of PyG 2.0.x's default rules:
  edge_index                      += cumulative N_0            (data_utils.py:30-31)
  hierarchy_edge_index_l          += cumulative N_l            (data_utils.py:36-37)
  hierarchy_trace_index_l         += cumulative N_l            (data_utils.py:39-40)
  hierarchy_dil_d_edge_index_l    += cumulative N_0  (falls through to PyG default
                                     `num_nodes`; data_utils.py:42 -- a reference quirk, wrong for B>1 and l>0)
  x / color / pos / mask / labels    concatenated on dim 0, no offset (data_utils.py:32-33)
  num_vertices                       stacked to [B, L+1] (`__cat_dim__` -> None, data_utils.py:23-25)
All index tensors stay int64, exactly as the reference delivers them.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List

import torch

_NO_OFFSET = ("x", "color", "pos", "mask", "labels")
_EDGE_L = re.compile(r"^hierarchy_edge_index_(\d+)$")
_TRACE_L = re.compile(r"^hierarchy_trace_index_(\d+)$")


class GraphBatch:
    """Attribute + string-key container (duck-typed PyG Data/Batch)."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    @property
    def keys(self) -> List[str]:
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __getitem__(self, key: str):
        try:
            return self.__dict__[key]
        except KeyError:
            raise KeyError(key) from None

    def __setitem__(self, key: str, value):
        setattr(self, key, value)

    def __contains__(self, key: str) -> bool:
        return key in self.__dict__ and self.__dict__[key] is not None

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def _host_num_vertices(self):
        """[B][L+1] vertex counts as python ints, taken while `num_vertices` is still a host tensor, so the device
        copy of the batch can be consumed without a device-to-host read."""
        pre = self.__dict__.get("_nv_host")
        nv = self.__dict__.get("num_vertices")
        if pre is None and torch.is_tensor(nv) and not nv.is_cuda:
            nv2 = nv if nv.dim() == 2 else nv.unsqueeze(0)
            pre = tuple(tuple(int(x) for x in g) for g in nv2.tolist())
        return pre

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        out = GraphBatch()
        for k in self.keys:
            v = self.__dict__[k]
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        out.__dict__["_nv_host"] = self._host_num_vertices()
        return out

    def pin_memory(self) -> "GraphBatch":
        out = GraphBatch()
        for k in self.keys:
            v = self.__dict__[k]
            out.__dict__[k] = v.pin_memory() if (torch.is_tensor(v) and not v.is_cuda) else v
        out.__dict__["_nv_host"] = self._host_num_vertices()
        return out

    def tensor_bytes(self, host_only: bool = False) -> int:
        return sum(v.numel() * v.element_size() for v in self.__dict__.values()
                   if torch.is_tensor(v) and not (host_only and v.is_cuda))


def _inc(sample: GraphBatch, key: str) -> int:
    nv = sample.num_vertices
    if key == "edge_index":
        return int(nv[0])
    if key in _NO_OFFSET:
        return 0
    m = _EDGE_L.match(key) or _TRACE_L.match(key)
    if m and 1 <= int(m.group(1)) < len(nv):
        return int(nv[int(m.group(1))])
    # PyG 2.0.x default Data.__inc__
    if "batch" in key:
        return int(sample[key].max()) + 1
    if "index" in key or "face" in key:
        return sample.num_nodes
    return 0


def _cat_dim(key: str):
    if key == "num_vertices":
        return None
    if "index" in key or "face" in key:
        return -1
    return 0


def collate(samples: Iterable[GraphBatch], keep_index: bool = True) -> GraphBatch:
    """Batch a list of single-graph samples the way PyG's DataLoader batches HierarchicalData.
    keep_index=False leaves the COO edge sets and trace maps out (their structure is cached per sample and batched
    on the device by stinet_b200.structure.attach_batch_structure)."""
    samples = list(samples)
    out = GraphBatch()
    fields: Dict[str, object] = {}
    for key in samples[0].keys:
        if not keep_index and (key == "edge_index" or key.startswith("hierarchy_")):
            continue
        vals, inc = [], 0
        for s in samples:
            v = s[key]
            if torch.is_tensor(v):
                step = _inc(s, key)
                if inc:
                    v = v + inc
                inc += step
            vals.append(v)
        if torch.is_tensor(vals[0]):
            cd = _cat_dim(key)
            fields[key] = torch.stack(vals, 0) if (cd is None or vals[0].dim() == 0) else torch.cat(vals, cd)
        else:
            fields[key] = vals
    n0 = torch.tensor([s.num_nodes for s in samples], dtype=torch.long)
    fields["batch"] = torch.repeat_interleave(torch.arange(len(samples)), n0)
    fields["ptr"] = torch.cat([torch.zeros(1, dtype=torch.long), n0.cumsum(0)])
    for k, v in fields.items():
        setattr(out, k, v)
    out.num_graphs = len(samples)
    return out
