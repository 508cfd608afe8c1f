"""``train_mm.py --module=<cogmen|dgcn|dgcnv2|mmgcn|dagerc> --dataset=... --modality=atv`` without lumo / accelerate
(reference: train_mm.py:16-25 -> track_mm/<module>.main -> mmbase.main, mmbase.py:483-499).

What is kept from the reference
  * the command line: ``--key=value`` pairs, the same parameter names (module, dataset, modality, epoch, seed, device,
    n_classes, train.batch_size ...), the per-module defaults of the ``*Params`` classes (cogmen.py:36-52, dgcn.py:24-44,
    mmgcn.py:22-50, dagerc.py:25-66) and the data-set -> feature-size rules of ``MMBaseParams.iparams`` (mmbase.py:52-128);
  * the train / test steps (cogmen.py:179-195, dgcn.py:117-134, mmgcn.py:141-157, dagerc.py:217-237, mmbase.py:152-201):
    mean cross entropy (class-weighted for 6-way DialogueGCN, masked for DAG-ERC + clip_grad_norm_(5)), Adam / AdamW with the
    reference's learning rates, a test pass after every epoch (EvalCallback(test_per_epoch=1), mmbase.py:136) with accuracy,
    weighted / micro / macro F1, balanced accuracy and the confusion matrix (sklearn, mmbase.py:262-283), seed marking per
    epoch, ``best_model.ckpt`` / ``last_model.ckpt`` as plain ``state_dict`` files (mmbase.py:325-333) whose keys are the
    reference modules' keys.
What is different
  * the model runs in libercgraph kernels; the batch is collated onto the device in the packed layout (collate.DeviceCollate);
    the optimizer is the flat-buffer Adam kernel (optim.FlatAdam);
  * ``mmdatasets`` (the reference's feature pickles) is not part of this repo: ``--data=<file>`` loads a pickle
    ``{"train": [sample, ...], "test": [...]}`` of samples in the reference's format (mmbase.py:374-384); without it a seeded
    synthetic data set of the data set's SHAPE is generated (``--train_dialogues``, ``--test_dialogues``).
"""
import json
import math
import os
import pickle
import sys
import time

import numpy as np
import torch

MODULES = ("cogmen", "dgcn", "dgcnv2", "mmgcn", "dagerc")


# ------------------------------------------------------------------------------------------------ parameters
def parse_args(argv):
    """fire / lumo style: ``--key=value``, ``--key value``, ``--flag`` (= True); dotted keys kept as given."""
    out, i = {}, 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("--"):
            raise SystemExit("unexpected argument %r (expected --key=value)" % a)
        if "=" in a:
            k, v = a[2:].split("=", 1)
        elif i + 1 < len(argv) and not argv[i + 1].startswith("--"):
            k, v = a[2:], argv[i + 1]
            i += 1
        else:
            k, v = a[2:], "True"
        out[k.replace("-", "_")] = _coerce(v)
        i += 1
    return out


def _coerce(v):
    if v in ("True", "true"):
        return True
    if v in ("False", "false"):
        return False
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def resolve_params(args):
    """Defaults of <Module>Params + MMBaseParams.iparams (mmbase.py:52-128)."""
    module = args.get("module")
    if module not in MODULES:
        raise SystemExit("--module= one of %s" % (MODULES,))
    p = dict(seed=1, dataset="iemocap-cogmen-6", modality="atv", n_speakers=2, device=0, reimplement=False, data=None,
             train_dialogues=120, test_dialogues=31, save_dir=None, log_every=0, max_len=None)
    per_module = {
        "cogmen": dict(epoch=55, batch_size=32, optim="Adam", lr=1e-4, weight_decay=1e-8, num_heads=17),          # cogmen.py:36-52
        "dgcn": dict(epoch=55, batch_size=32, optim="Adam", lr=3e-4, weight_decay=0.0, loss_weights=True),        # dgcn.py:24-44
        "dgcnv2": dict(epoch=55, batch_size=32, optim="Adam", lr=3e-4, weight_decay=0.0, loss_weights=True, base_model="LSTM"),   # dgcnv2.py:20-45
        "mmgcn": dict(epoch=60, batch_size=16, optim="Adam", lr=3e-4, weight_decay=3e-5),                          # mmgcn.py:22-40
        "dagerc": dict(epoch=30, batch_size=8, optim="AdamW", lr=1e-3, weight_decay=1e-2, gnn_layers=4, dropout=0.0),   # dagerc.py:25-41
    }[module]
    p.update(per_module)
    p["module"] = module
    for k, v in args.items():
        k = {"train.batch_size": "batch_size", "optim.lr": "lr", "optim.weight_decay": "weight_decay"}.get(k, k)
        p[k] = v
    ds = p["dataset"]
    if "n_classes" not in p:
        tail = ds.rsplit("-", 1)[-1]
        p["n_classes"] = int(tail) if tail.isdigit() else 6
    ht, ha, hv = 100, 100, 100                                                                    # mmbase.py:45-47
    if "iemocap" in ds and "cogmen" in ds:
        ha, ht, hv = 100, 100, 512
    elif "meld" in ds:
        p["n_speakers"] = 9
        if "mmgcn" in ds:
            ha, ht, hv = 300, 600, 342
    elif "mosei" in ds:
        ht, ha, hv = 300, 74, 35
    if "pad80" in ds:
        ha = 80
    elif "fbank" in ds:
        ha = 640
    elif "is10" in ds:
        ha = 1584
    if "sbert" in ds or "robert" in ds:
        ht = 768
    if "tsn" in ds:
        hv = hv + 2048 if "v+" in ds else 2048
    p.update(hidden_text=ht, hidden_audio=ha, hidden_visual=hv)
    p["hidden_all"] = sum({"t": ht, "a": ha, "v": hv}[m] for m in set(p["modality"]))
    if module == "dagerc" and p["reimplement"] and "iemocap" in ds:                               # dagerc.py:45-50
        p.update(dropout=0.2, epoch=55, batch_size=16, lr=5e-4)
    p["batch_first"] = module not in ("mmgcn", "dgcnv2")                                          # mmgcn.py:39-40, dgcnv2.py:45
    p["speaker_onehot"] = module in ("mmgcn", "dagerc", "dgcnv2")                                 # mmgcn.py:39, dagerc.py:41, dgcnv2.py:44
    return p


# ------------------------------------------------------------------------------------------------ data
def synthetic_dataset(n_dialogues, p, seed, one_speaker=False):
    """Dialogues in the reference's sample format with the data set's feature sizes; IEMOCAP-like lengths
    (clamp(round(49 + 21 z), 8, 110)) or MOSEI-like (1 + Geometric(1/7) <= 40, one speaker id, mosei_feature.py:211)."""
    rng = np.random.default_rng(seed)
    mosei = "mosei" in p["dataset"]
    out = []
    for _ in range(n_dialogues):
        if mosei:
            L = int(min(40, rng.geometric(1.0 / 7.0)))
        else:
            L = int(np.clip(round(49 + 21 * rng.standard_normal()), 8, 110))
        if p.get("max_len"):
            L = min(L, int(p["max_len"]))
        spk = np.zeros(L, dtype=np.int64) if (mosei or one_speaker) else rng.integers(0, p["n_speakers"], size=L)
        label = rng.integers(0, p["n_classes"], size=L)
        # a learnable signal: the class shifts the first few text / audio dimensions
        t = rng.standard_normal((L, p["hidden_text"])).astype(np.float32)
        a = rng.standard_normal((L, p["hidden_audio"])).astype(np.float32)
        v = rng.standard_normal((L, p["hidden_visual"])).astype(np.float32)
        t[np.arange(L), label % p["hidden_text"]] += 3.0
        a[np.arange(L), label % p["hidden_audio"]] += 3.0
        out.append(({"text": list(t), "audio": list(a), "visual": list(v),
                     "speakers": [[1 if s == k else 0 for k in range(p["n_speakers"])] for s in spk],
                     "label": [int(x) for x in label]},))
    return out


def load_data(p):
    if p.get("data"):
        with open(p["data"], "rb") as f:
            d = pickle.load(f)
        wrap = lambda xs: [s if isinstance(s, tuple) else (s,) for s in xs]
        return wrap(d["train"]), wrap(d["test"]), "file %s" % p["data"]
    return (synthetic_dataset(p["train_dialogues"], p, 1000 + p["seed"]), synthetic_dataset(p["test_dialogues"], p, 2000 + p["seed"]),
            "synthetic (%s shape: t=%d a=%d v=%d)" % (p["dataset"], p["hidden_text"], p["hidden_audio"], p["hidden_visual"]))


# ------------------------------------------------------------------------------------------------ trainer
DGCN_LOSS_WEIGHTS = [1 / 0.086747, 1 / 0.144406, 1 / 0.227883, 1 / 0.160585, 1 / 0.127711, 1 / 0.252668]      # dgcn.py:109-110


def build_model(p, dev):
    if p["module"] == "cogmen":
        from .track_mm.cogmen import COGMENModule
        m = COGMENModule(input_size=p["hidden_all"], hidden_size=100, num_head=p["num_heads"], n_speakers=p["n_speakers"],
                         n_classes=p["n_classes"], build_dead_encoder=bool(p.get("dead_encoder", False)))
    elif p["module"] == "dgcn":
        from .track_mm.dgcn import DGCNModule
        m = DGCNModule(input_size=p["hidden_all"], hidden_size=200, n_speakers=p["n_speakers"], n_classes=p["n_classes"])
    elif p["module"] == "dgcnv2":
        from .track_mm.dgcnv2 import DGCNModule as DGCNv2Module
        m = DGCNv2Module(base_model=p["base_model"], input_size=p["hidden_all"], hidden_size=100, n_speakers=p["n_speakers"],
                         n_classes=p["n_classes"], context_attention="general")                      # dgcnv2.py:188-195
    elif p["module"] == "mmgcn":
        from .track_mm.mmgcn import MMGCNModule
        m = MMGCNModule(hidden_text=p["hidden_text"], hidden_visual=p["hidden_visual"], hidden_audio=p["hidden_audio"],
                        n_speakers=p["n_speakers"], n_classes=p["n_classes"], modals=p["modality"])
    else:
        from .track_mm.dagerc import DAGERCModule
        m = DAGERCModule(emb_dim=p["hidden_all"], dropout=p["dropout"], gnn_layers=p["gnn_layers"], n_classes=p["n_classes"])
    return m.to(dev)


class Trainer:
    def __init__(self, p):
        from .collate import DeviceCollate
        from .optim import FlatAdam
        self.p = p
        self.dev = torch.device("cuda", int(p["device"])) if p["device"] != "cpu" else None
        if self.dev is None or not torch.cuda.is_available():
            raise SystemExit("train_mm: the libercgraph modules need a CUDA device (there is no CPU path)")
        torch.cuda.set_device(self.dev)
        torch.manual_seed(p["seed"])
        self.model = build_model(p, self.dev)
        self.collate = DeviceCollate(p["modality"], p["batch_first"], p["speaker_onehot"], p["n_speakers"], device=self.dev)
        self.optim = FlatAdam(self.model.parameters(), lr=p["lr"], weight_decay=p["weight_decay"], decoupled=p["optim"] == "AdamW",
                              max_norm=5.0 if p["module"] == "dagerc" else None)                    # dagerc.py:229-231
        self.class_weight = None
        if p["module"] in ("dgcn", "dgcnv2") and p.get("loss_weights") and p["n_classes"] == 6:
            self.class_weight = torch.tensor(DGCN_LOSS_WEIGHTS, device=self.dev)
        self.best = {}

    def _logits(self, batch):
        m = self.p["module"]
        if m == "cogmen":
            return self.model(**batch.packed_kwargs())[0]
        if m == "dgcn":
            return self.model(input_tensor=batch["input_tensor"], speaker_tensor=batch["speaker_tensor"], text_length=batch["text_length"])[0]
        if m == "dgcnv2":
            return self.model(input_tensor=batch["input_tensor"], speaker_tensor=batch["speaker_tensor"],
                              attention_mask=batch["attention_mask"], text_length=batch["text_length"])[0]
        if m == "mmgcn":
            return self.model(text_feature=batch["text_feature"], audio_feature=batch["audio_feature"],
                              visual_feature=batch["visual_feature"], speaker_tensor=batch["speaker_tensor"],
                              text_length=batch["text_length"])[0]
        logits = self.model(input_tensor=batch["input_tensor"], text_length=batch["text_length"], speaker_tensor=batch["speaker_tensor"])[0]
        return logits[batch["attention_mask"].bool()]                                               # dagerc.py:223-226

    def train_step(self, samples):
        from . import ops
        batch = self.collate(samples)
        logits = self._logits(batch)
        loss = ops.cross_entropy(logits, batch["label"], self.class_weight)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        with torch.no_grad():
            acc = (logits.argmax(-1) == batch["label"]).float().mean()
        return loss.detach(), acc

    @torch.no_grad()
    def test(self, data):
        """mmbase.py:180-201 + on_test_end :262-283."""
        from sklearn import metrics
        from . import ops
        self.model.eval()
        true, pred, lsum, n = [], [], 0.0, 0
        bs = self.p["batch_size"]
        for i in range(0, len(data), bs):
            batch = self.collate(data[i:i + bs])
            logits = self._logits(batch)
            ys = batch["label"]
            lsum += float(ops.cross_entropy(logits, ys)) * ys.numel()
            n += ys.numel()
            true.extend(ys.cpu().tolist())
            pred.extend(logits.argmax(-1).cpu().tolist())
        self.model.train()
        return {"Lall": lsum / max(n, 1), "acc": metrics.accuracy_score(true, pred),
                "wa": metrics.balanced_accuracy_score(true, pred),
                "f1": metrics.f1_score(true, pred, average="weighted"), "mif1": metrics.f1_score(true, pred, average="micro"),
                "maf1": metrics.f1_score(true, pred, average="macro"),
                "cm": metrics.confusion_matrix(true, pred, labels=range(self.p["n_classes"])).tolist(), "C": n}

    def save(self, name):
        d = self.p.get("save_dir")
        if not d:
            return None
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, name)
        torch.save({k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}, path)   # mmbase.py:325-333
        return path

    def train(self, train, test, log=print):
        p = self.p
        self.model.train()
        history = []
        for epoch in range(1, int(p["epoch"]) + 1):
            rng = np.random.default_rng(p["seed"] * 100003 + epoch)                                  # rnd.mark(seed) per epoch, mmbase.py:207-208
            order = rng.permutation(len(train))
            t0 = time.perf_counter()
            lsum, asum, nb, utts = 0.0, 0.0, 0, 0
            for i in range(0, len(order), p["batch_size"]):
                samples = [train[j] for j in order[i:i + p["batch_size"]]]
                loss, acc = self.train_step(samples)
                lsum += float(loss)
                asum += float(acc)
                nb += 1
                utts += sum(len(s[0]["text"]) for s in samples)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            te = self.test(test)                                                                      # test_per_epoch = 1, mmbase.py:136
            rec = {"epoch": epoch, "train_Lall": lsum / nb, "train_Acc": asum / nb, "utt_per_s": utts / dt,
                   **{"test_" + k: v for k, v in te.items() if k != "cm"}}
            history.append(rec)
            if te["f1"] >= self.best.get("f1", -1.0):
                self.best = {"f1": te["f1"], "acc": te["acc"], "epoch": epoch, "cm": te["cm"]}
                self.save("best_model.ckpt")
            log("epoch %3d  train Lall %.4f Acc %.3f  |  test Lall %.4f acc %.3f f1 %.3f  |  %.0f utt/s" % (
                epoch, rec["train_Lall"], rec["train_Acc"], te["Lall"], te["acc"], te["f1"], rec["utt_per_s"]))
        self.save("last_model.ckpt")
        return history


def main(argv=None, log=print):
    args = parse_args(list(sys.argv[1:] if argv is None else argv))
    p = resolve_params(args)
    train, test, what = load_data(p)
    log("train_mm: module=%s dataset=%s modality=%s hidden_all=%d n_classes=%d n_speakers=%d | %s: %d / %d dialogues" % (
        p["module"], p["dataset"], p["modality"], p["hidden_all"], p["n_classes"], p["n_speakers"], what, len(train), len(test)))
    tr = Trainer(p)
    if p.get("eval_first"):
        log("eval_first:", {k: v for k, v in tr.test(test).items() if k != "cm"})
    hist = tr.train(train, test, log=log)
    log("Best Results", json.dumps({k: v for k, v in tr.best.items() if k != "cm"}))
    return {"params": p, "history": hist, "best": tr.best, "trainer": tr}


if __name__ == "__main__":
    main()
