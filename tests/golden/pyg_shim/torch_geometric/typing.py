from typing import Optional, Tuple, Union  # noqa: F401  (reference does `from torch_geometric.typing import Union, Tuple`)
from torch import Tensor

Adj = Tensor
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
Size = Optional[Tuple[int, int]]
