"""INTEGRATION.md "Level 1", executed: the REAL reference model files with only their import lines redirected to this
package -- torch_geometric.nn -> erc_b200.pyg_nn, models.rgcn -> erc_b200.models.rgcn, cogmen_utils / dgcn_models
batch_graphify -> erc_b200's.  Run as a script in its own interpreter (it rewires sys.modules):

    python tests/level1_swap_check.py            construct + load reference-shaped state_dicts (CPU box)
                                                 (+ one forward/backward of each when a GPU and /root/reference are both present)
Prints one JSON line.  TEST INFRASTRUCTURE: needs /root/reference, so it only runs in the build container."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from oracle import ref_loader, pyg_standin
    if not ref_loader.available():
        print(json.dumps({"skipped": "reference tree not present"}))
        return 0
    # 1. what the UNMODIFIED reference produces (stand-in PyG layers): state_dict keys / shapes to be loaded below
    ref = ref_loader.load()
    torch.manual_seed(0)
    want_cogmen = {k: v.clone() for k, v in ref.cogmen.COGMENModule(1380, 100, 17, 2, 4).state_dict().items()}
    want_dgcn = {k: v.clone() for k, v in ref.dgcn.DGCNModule(2, input_size=1380, hidden_size=200, n_classes=6).state_dict().items()}
    # 2. the import swaps of INTEGRATION.md Level 1
    import erc_b200  # noqa: F401
    from erc_b200 import pyg_nn
    from erc_b200.models import rgcn as our_rgcn
    from erc_b200.track_mm import cogmen_utils as our_cu, dgcn_models as our_dm
    for name in [n for n in sys.modules if n.startswith("track_mm.") and n.split(".")[1] in ("cogmen", "dgcn", "dgcn_models", "cogmen_utils")]:
        del sys.modules[name]
    sys.modules.pop("models.rgcn", None)
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_nn.RGCNConv, tg_nn.TransformerConv, tg_nn.GraphConv = pyg_nn.RGCNConv, pyg_nn.TransformerConv, pyg_nn.GraphConv
    sys.modules["torch_geometric.nn"] = tg_nn
    sys.modules["torch_geometric"].nn = tg_nn
    sys.modules["models.rgcn"] = our_rgcn                      # dgcn_models.py:7  from models.rgcn import RGCNConv
    import importlib
    ref_cu = importlib.import_module("track_mm.cogmen_utils")
    ref_cu.batch_graphify = our_cu.batch_graphify             # cogmen.py:32   from .cogmen_utils import batch_graphify
    ref_cogmen = importlib.import_module("track_mm.cogmen")
    ref_dm = importlib.import_module("track_mm.dgcn_models")
    for fn in ("batch_graphify", "SeqContext", "EdgeAtt"):     # dgcn.py:20     from .dgcn_models import (...)
        setattr(ref_dm, fn, getattr(our_dm, fn))
    ref_dgcn = importlib.import_module("track_mm.dgcn")
    out = {}
    m = ref_cogmen.COGMENModule(1380, 100, 17, 2, 4)          # the reference's OWN class, our layers inside
    assert type(m.gcn.conv1) is pyg_nn.RGCNConv and type(m.gcn.conv2) is pyg_nn.TransformerConv
    m.load_state_dict(want_cogmen, strict=True)
    out["cogmen_keys"] = len(want_cogmen)
    d = ref_dgcn.DGCNModule(2, input_size=1380, hidden_size=200, n_classes=6)
    assert type(d.gcn.conv1) is our_rgcn.RGCNConv and type(d.gcn.conv2) is pyg_nn.GraphConv
    d.load_state_dict(want_dgcn, strict=True)
    out["dgcn_keys"] = len(want_dgcn)
    if torch.cuda.is_available():                             # never true where /root/reference exists today; kept for a box that has both
        from erc_b200 import synth
        b = synth.config1(seed=0, B=4)
        m = m.cuda().eval()
        with torch.no_grad():
            logits, _ = m(b["input_tensor"].cuda(), b["speaker_tensor"].cuda(), b["text_length"])
        out["cogmen_forward"] = list(logits.shape)
    else:
        from erc_b200 import synth
        b = synth.config1(seed=0, B=2)
        try:
            m.eval()
            m(b["input_tensor"], b["speaker_tensor"], b["text_length"])
            out["cpu_forward"] = "ran (unexpected: there is no CPU path)"
        except Exception as e:                                # the product has no CPU fallback: it must refuse, loudly
            out["cpu_forward_refused"] = type(e).__name__
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
