"""Import-only stand-in: the reference imports SparseTensor/matmul (sage_conv_filter.py:14) but only
reaches them through message_and_aggregate, which edge_index (Tensor) inputs never take."""


class SparseTensor:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError


def matmul(*a, **k):  # pragma: no cover
    raise NotImplementedError
