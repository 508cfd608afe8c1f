"""B200-native conversation-graph hot path of sailist/emotion-recognition-in-conversation.

Import name: the directory is called ``emotion-recognition-in-conversation_b200`` (not a Python
identifier); ``import erc_b200`` (repo-root shim) loads it under that alias.
"""
from . import _lib, ops, graph, pyg_nn  # noqa: F401
from ._lib import ErcgError, launch_count  # noqa: F401
from .graph import PackedGraph, build_graph, graph_from_edge_index  # noqa: F401
