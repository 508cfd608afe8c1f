"""torch.autograd wrapper for K11 (MaskedEdgeAttention 'attn1' in closed form, csrc/dgcnv2.cu; include/ercgraph.h)."""
import torch

from ._lib import lib, check
from .ops import _p, _stream, _rows


class _MaskedEdgeAtt(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, graph, Lmax, wp, wf):
        S, ldS = _rows(S)
        N = graph.N
        nu = torch.empty(graph.E, dtype=torch.float32, device=S.device)
        stat = torch.empty((N, 2), dtype=torch.float32, device=S.device)
        check(lib().ercg_masked_edge_att_fwd(_p(S), ldS, _p(graph.node_off), _p(graph.node_dlg), _p(graph.t_rowptr), _p(graph.t_eid),
                                             int(Lmax), int(wp), int(wf), _p(nu), _p(stat), N, _stream()), "ercg_masked_edge_att_fwd")
        ctx.graph, ctx.Lmax, ctx.wp, ctx.wf = graph, int(Lmax), int(wp), int(wf)
        ctx.save_for_backward(S, nu, stat)
        return nu

    @staticmethod
    def backward(ctx, dnu):
        S, nu, stat = ctx.saved_tensors
        g = ctx.graph
        dS = torch.zeros_like(S)                       # columns of positions beyond a dialogue's length stay zero
        check(lib().ercg_masked_edge_att_bwd(_p(S), S.stride(0), _p(g.node_off), _p(g.node_dlg), _p(g.t_rowptr), _p(g.t_eid),
                                             ctx.Lmax, ctx.wp, ctx.wf, _p(nu), _p(dnu.contiguous()), _p(stat), _p(dS), dS.stride(0),
                                             g.N, _stream()), "ercg_masked_edge_att_bwd")
        return dS, None, None, None, None


def masked_edge_att(S, graph, Lmax, wp, wf):
    """S [B*Lmax, >= Lmax'] (dialogue-major full-length rows x max_seq_len columns) -> edge weights in by-destination order."""
    assert S.size(0) == graph.B * int(Lmax) and S.size(1) >= int(Lmax) or S.size(1) >= 1
    return _MaskedEdgeAtt.apply(S, graph, Lmax, wp, wf)
