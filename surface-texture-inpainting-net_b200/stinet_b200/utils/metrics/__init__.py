from . import graph_metrics  # noqa: F401
