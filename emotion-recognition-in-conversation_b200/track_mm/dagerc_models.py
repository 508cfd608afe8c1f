"""Drop-in for the live classes of track_mm/dagerc_models.py: ``GAT_dialoggcn_v1`` (:312-365), ``mask_logic`` (:83-90)
and ``attentive_node_features`` (:425-, identity for nodal_att_type=None which is all DAGERCModule uses).

In the reference the GAT is called once per utterance and layer from a Python loop (dagerc.py:174); here the whole loop
is the persistent kernel K10 (csrc/dagerc.cu) driven by ``DAGERCModule``.  ``GAT_dialoggcn_v1`` keeps its parameters
(``linear``, ``Wr0``, ``Wr1``: same state_dict keys) and K10 reads them; called on its own, ``forward(Q, K, V, adj, s_mask)``
runs the stand-alone step kernel (csrc/dag_gat.cu)."""
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
        """Stand-alone step with the reference's signature (dagerc_models.py:326-365):
        Q [B,D], K / V [B,N,D], adj [B,N], s_mask [B,N]  ->  (attn_weight [B,1,N], attn_sum [B,D]).
        DAGERCModule does not come through here (its utterance loop is the fused layer kernel K10); this is for callers
        that use the class on its own.  CUDA only: scores + mask + softmax + the two speaker-selected weighted sums are
        one kernel (ercg_dag_gat_fwd), the Wr0 / Wr1 mix is one dense transform on the [B,2D] result."""
        from .. import ops, ops_dagerc
        alpha, s01 = ops_dagerc.gat_step(Q, K, V, adj, s_mask, self.linear.weight, self.linear.bias)
        attn_sum = ops.linear(s01, torch.cat([self.Wr0.weight, self.Wr1.weight], 1))
        return alpha.unsqueeze(1), attn_sum


class attentive_node_features(nn.Module):
    """Dead in the reference's forward (nodal_att_type is None, dagerc.py:84,193); kept for its state_dict keys."""

    def __init__(self, hidden_size):
        super().__init__()
        self.transform = nn.Linear(hidden_size, hidden_size)

    def forward(self, features, lengths, nodal_att_type):
        if nodal_att_type is None:
            return features
        raise NotImplementedError("DAGERCModule only uses nodal_att_type=None (track_mm/dagerc.py:84)")
