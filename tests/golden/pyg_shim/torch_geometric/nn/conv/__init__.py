from .message_passing import MessagePassing
from .edge_conv import EdgeConv
from .sage_conv import SAGEConv
from . import edge_conv, sage_conv, message_passing  # noqa: F401
