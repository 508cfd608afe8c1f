"""Drop-in for the live classes of track_mm/dagerc_models.py: ``GAT_dialoggcn_v1`` (:312-365), ``mask_logic`` (:83-90)
and ``attentive_node_features`` (:425-, identity for nodal_att_type=None which is all DAGERCModule uses).

In the reference the GAT is called once per utterance and layer from a Python loop (dagerc.py:174); here the whole loop
is the persistent kernel K10 (csrc/dagerc.cu) driven by ``DAGERCModule``.  ``GAT_dialoggcn_v1`` keeps its parameters
(``linear``, ``Wr0``, ``Wr1``: same state_dict keys) and K10 reads them."""
import torch
from torch import nn


def mask_logic(alpha, adj):
    return alpha - (1 - adj) * 1e30


class GAT_dialoggcn_v1(nn.Module):
    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.linear = nn.Linear(hidden_size * 2, 1)
        self.Wr0 = nn.Linear(hidden_size, hidden_size, bias=False)
        self.Wr1 = nn.Linear(hidden_size, hidden_size, bias=False)

    def forward(self, Q, K, V, adj, s_mask):
        raise NotImplementedError(
            "the per-utterance GAT step is fused into the persistent DAG layer kernel (erc_b200.ops_dagerc.dag_layer); "
            "call DAGERCModule.forward, there is no stand-alone CPU/ATen path")


class attentive_node_features(nn.Module):
    """Dead in the reference's forward (nodal_att_type is None, dagerc.py:84,193); kept for its state_dict keys."""

    def __init__(self, hidden_size):
        super().__init__()
        self.transform = nn.Linear(hidden_size, hidden_size)

    def forward(self, features, lengths, nodal_att_type):
        if nodal_att_type is None:
            return features
        raise NotImplementedError("DAGERCModule only uses nodal_att_type=None (track_mm/dagerc.py:84)")
