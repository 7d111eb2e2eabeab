"""Mirror of reference models/modules/__init__.py:6-10, which re-exports every class of every module."""
from . import (edge_conv_filter, edge_conv_translation_invariance, fastinstancenorm, sage_conv_filter,  # noqa: F401
               singlebatchgroupnorm)
from .edge_conv_filter import EdgeConv  # noqa: F401
from .edge_conv_translation_invariance import EdgeConvTransInv  # noqa: F401
from .fastinstancenorm import FastInstanceNorm  # noqa: F401
from .sage_conv_filter import SAGEConv, SAGEConvTransInv  # noqa: F401
from .singlebatchgroupnorm import SingleBatchGraphNorm  # noqa: F401
