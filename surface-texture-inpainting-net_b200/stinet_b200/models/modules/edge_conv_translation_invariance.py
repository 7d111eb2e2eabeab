"""Mirror of reference models/modules/edge_conv_translation_invariance.py:9-24: the first conv of the 3D network
uses nn(x_j - x_i) so absolute positions never enter the message."""
from .edge_conv_filter import EdgeConv


class EdgeConvTransInv(EdgeConv):
    trans_inv = True

    def __init__(self, nn, aggr):
        super().__init__(nn, aggr)
