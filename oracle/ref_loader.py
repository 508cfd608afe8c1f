"""Import the *real* reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``-m "not gpu"`` tests that cross-check the restatement against the live
reference when the tree is present.  Nothing under ``-m gpu``, ``smoke()`` or
``bench.py`` may call this: /root/reference does not exist on the GPU box.

Recipe: the reference's trainer stack (lumo, accelerate,
fire, omegaconf, dbrecord, h5py) and torch_geometric / torch_scatter are not
installed, so before importing ``track_mm.{cogmen,dgcn,mmgcn,dagerc}`` we
pre-seed ``sys.modules`` with
  (1) ``torch_scatter.scatter_add``             -> oracle.pyg_standin.scatter_add
  (2) ``torch_geometric.nn.{RGCNConv,TransformerConv,GraphConv}`` -> oracle.pyg_standin
  (3) inert placeholders for lumo*, mmdatasets*, track_mm.mmbase
and apply the torch-2.11 / numpy-2 compatibility patches listed in SURVEY.md
section 8c (semantics unchanged):
  (i)  contrib/nn.py:268  TransformerEncoderLayer.forward gains an ignored
       ``is_causal`` keyword (nn.TransformerEncoder passes it since torch 2.0);
  (ii) track_mm/mmgcn_models.py:634  ``adj[idx] = dia_sim`` with an ndarray
       ``idx`` of shape [2,L] is read as ``adj[idx[0], idx[1]] = dia_sim``
       (old numpy "ndarray as tuple of index arrays" behaviour).
"""
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ERC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "track_mm", "cogmen_utils.py"))


def _placeholder(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Inert:
    """Base class stand-in for lumo/mmbase classes the model files subclass at import time."""

    def __init__(self, *a, **k):
        pass


_loaded = {}


def load():
    """Returns a namespace with the reference's own classes/functions."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    from . import pyg_standin

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    _placeholder("torch_scatter", scatter_add=pyg_standin.scatter_add)
    tg = _placeholder("torch_geometric")
    tg.nn = _placeholder("torch_geometric.nn", RGCNConv=pyg_standin.RGCNConv,
                         TransformerConv=pyg_standin.TransformerConv, GraphConv=pyg_standin.GraphConv)
    lumo = _placeholder("lumo", CollateBase=_Inert, Meter=_Inert, DataModule=_Inert, MetricType=dict,
                        TrainStage=_Inert, callbacks=types.SimpleNamespace(), Trainer=_Inert, Record=_Inert)
    lumo.contrib = _placeholder("lumo.contrib", EMA=_Inert)
    _placeholder("mmdatasets")
    _placeholder("mmdatasets.erc_dataset", get_train_dataset=None, get_test_dataset=None, get_val_dataset=None)

    track_mm = importlib.import_module("track_mm")
    _placeholder("track_mm.mmbase", MMBaseTrainer=_Inert, MMBaseParams=_Inert, main=lambda *a, **k: None)

    # (i) is_causal compat patch
    cnn = importlib.import_module("contrib.nn")
    if not getattr(cnn.TransformerEncoderLayer.forward, "_erc_patched", False):
        _orig = cnn.TransformerEncoderLayer.forward

        def _fwd(self, src, src_mask=None, src_key_padding_mask=None, is_causal=False):
            return _orig(self, src, src_mask, src_key_padding_mask)

        _fwd._erc_patched = True
        cnn.TransformerEncoderLayer.forward = _fwd

    # (ii) numpy-2 indexing compat patch: exec a patched copy of mmgcn_models under its own name
    path = os.path.join(REF_ROOT, "track_mm", "mmgcn_models.py")
    with open(path) as f:
        text = f.read()
    assert "adj[idx] = dia_sim" in text
    text = text.replace("adj[idx] = dia_sim", "adj[idx[0], idx[1]] = dia_sim")
    spec = importlib.util.spec_from_loader("track_mm.mmgcn_models", loader=None, origin=path)
    mm_models = importlib.util.module_from_spec(spec)
    mm_models.__file__ = path
    mm_models.__package__ = "track_mm"
    sys.modules["track_mm.mmgcn_models"] = mm_models
    exec(compile(text, path, "exec"), mm_models.__dict__)
    track_mm.mmgcn_models = mm_models

    mods = {}
    for name in ("cogmen_utils", "cogmen", "dgcn_models", "dgcn", "mmgcn_utils", "mmgcn",
                 "dagerc_models", "dagerc"):
        mods[name] = importlib.import_module("track_mm." + name)
    mods["mmgcn_models"] = mm_models
    mods["rgcn"] = importlib.import_module("models.rgcn")
    _loaded.update(mods)
    return types.SimpleNamespace(**_loaded)


def load_dgcnv2():
    """The declare-lab DialogueGCN variant (track_mm/dgcnv2_models.py, dgcnv2.py).  Compat patch (iii) of SURVEY.md 8c:
    ``mask[edge_ind_] = 1`` with an ndarray ``edge_ind_`` of shape [3,E] (dgcnv2_models.py:557-559) relied on old numpy /
    torch treating the ndarray as a tuple of index arrays; it is read as ``mask[tuple(edge_ind_)] = 1``."""
    import re
    ns = load()
    if "dgcnv2" in _loaded:
        return types.SimpleNamespace(**_loaded)
    lumo = sys.modules["lumo"]
    contrib = sys.modules["lumo.contrib"]
    torch_pkg = _placeholder("lumo.contrib.torch")
    import torch

    def onehot(labels, label_num):
        return torch.zeros(*labels.shape, label_num, device=labels.device).scatter_(-1, labels.unsqueeze(-1), 1)

    tensor_mod = _placeholder("lumo.contrib.torch.tensor", onehot=onehot)
    contrib.torch = torch_pkg
    torch_pkg.tensor = tensor_mod
    path = os.path.join(REF_ROOT, "track_mm", "dgcnv2_models.py")
    with open(path) as f:
        text = f.read()
    assert "mask[edge_ind_] = 1" in text and "mask_copy[edge_ind_] = 1" in text
    text = text.replace("mask[edge_ind_] = 1", "mask[tuple(edge_ind_)] = 1").replace("mask_copy[edge_ind_] = 1", "mask_copy[tuple(edge_ind_)] = 1")
    spec = importlib.util.spec_from_loader("track_mm.dgcnv2_models", loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    mod.__package__ = "track_mm"
    sys.modules["track_mm.dgcnv2_models"] = mod
    exec(compile(text, path, "exec"), mod.__dict__)
    _loaded["dgcnv2_models"] = mod
    _loaded["dgcnv2"] = importlib.import_module("track_mm.dgcnv2")
    return types.SimpleNamespace(**_loaded)


def load_collate():
    """The reference's own ``ERCCollate`` class (track_mm/mmbase.py:344-455).  mmbase.py itself cannot be imported (lumo,
    dbrecord, mmdatasets are not installed), so the class statement is cut out of the file verbatim and executed with the
    two names it needs from lumo: ``CollateBase`` (a do-nothing base) and ``onehot`` (lumo/contrib/torch/tensor.py:57-67,
    restated: zeros(..., n).scatter_(-1, labels.unsqueeze(-1), 1))."""
    import re
    import torch
    path = os.path.join(REF_ROOT, "track_mm", "mmbase.py")
    with open(path) as f:
        text = f.read()
    m = re.search(r"^class ERCCollate\(CollateBase\):.*?(?=^class ERCDM)", text, flags=re.S | re.M)
    assert m, "ERCCollate not found in the reference"

    class CollateBase:
        def __init__(self, params=None):
            pass

    def onehot(labels, label_num):
        return torch.zeros(*labels.shape, label_num, device=labels.device).scatter_(-1, labels.unsqueeze(-1), 1)

    ns = {"CollateBase": CollateBase, "onehot": onehot, "torch": torch, "ParamsType": object}
    exec(compile(m.group(0), path, "exec"), ns)
    return ns["ERCCollate"]
