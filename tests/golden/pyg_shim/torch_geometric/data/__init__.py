"""Data / Batch.from_data_list restated from PyG 2.0.x collate semantics (test-only):
for every key, values are concatenated along __cat_dim__ (None => stacked on a new dim 0) after adding the
running sum of __inc__ of the preceding samples; `batch` / `ptr` are derived from num_nodes."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        for k, v in kwargs.items():
            setattr(self, k, v)

    # -- storage -----------------------------------------------------------------
    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    @property
    def num_nodes(self):
        if getattr(self, "x", None) is not None:
            return self.x.size(0)
        if getattr(self, "pos", None) is not None:
            return self.pos.size(0)
        raise ValueError("cannot infer num_nodes")

    def to(self, device, *a, **k):
        for key in self.keys:
            v = self[key]
            if torch.is_tensor(v):
                self[key] = v.to(device, *a, **k)
        return self

    # -- collate rules (PyG 2.0.x data.py) ------------------------------------------
    def __cat_dim__(self, key, value, *args, **kwargs):
        if "index" in key or "face" in key:
            return -1
        return 0

    def __inc__(self, key, value, *args, **kwargs):
        if "batch" in key:
            return int(value.max()) + 1
        if "index" in key or "face" in key:
            return self.num_nodes
        return 0


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        first = data_list[0]
        for key in first.keys:
            vals, inc = [], 0
            for d in data_list:
                v = d[key]
                if torch.is_tensor(v):
                    step = d.__inc__(key, v)
                    if torch.is_tensor(step) or step != 0:
                        v = v + inc
                    inc = inc + step
                vals.append(v)
            v0 = vals[0]
            if torch.is_tensor(v0):
                cat_dim = first.__cat_dim__(key, v0)
                if cat_dim is None:
                    out[key] = torch.stack(vals, dim=0)
                elif v0.dim() == 0:
                    out[key] = torch.stack(vals, dim=0)
                else:
                    out[key] = torch.cat(vals, dim=cat_dim)
            elif isinstance(v0, (int, float)):
                out[key] = torch.tensor(vals)
            else:
                out[key] = vals
        n = [d.num_nodes for d in data_list]
        out.batch = torch.repeat_interleave(torch.arange(len(n)), torch.tensor(n))
        out.ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.tensor(n).cumsum(0)])
        out.num_graphs = len(n)
        return out
