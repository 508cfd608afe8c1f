"""ERCCollate (track_mm/mmbase.py:344-455) with the batch built ON THE DEVICE in the packed layout.

The reference collate runs on the CPU and, per dialogue, stacks every modality, zero-pads it to the batch's longest dialogue,
builds the concatenated ``input_tensor`` the same way (three redundant copies of every feature) and hands pageable tensors to
``accelerate`` for a synchronous H2D copy.  Here the host does ONE pass: every utterance's modality vectors are written side
by side (in the order of the letters of ``--modality``, mmbase.py:408-410) into a pinned ``[N, ld]`` staging buffer -- packed,
dialogue-major, row pitch a multiple of 4 floats -- and that buffer, the speaker ids and the labels cross PCIe (no padding
bytes).  Everything else the reference's batch dict carries is produced on the GPU from it:

  x_packed [N, hidden_all], speaker_packed [N], text_length [B]    what COGMENModule.forward / forward_packed consume directly
  input_tensor, text_/audio_/visual_feature (zero padded)          ercg_unpack_rows on column slices of x_packed (lazy: built on
                                                                    first access, so a model that takes the packed entry point
                                                                    never pays for them)
  attention_mask, speaker_tensor (ids, transposed, one-hot)        ercg_collate_masks
Same keys, shapes, dtypes and values as the reference (tests/test_collate.py pins them against the real ERCCollate).
"""
from collections.abc import Mapping

import numpy as np
import torch

from ._lib import lib, check
from . import ops
from .ops import _p, _stream


class _LazyBatch(Mapping):
    """dict-like batch: padded tensors are materialised on first access (``batch['input_tensor']``, ``**batch``)."""

    def __init__(self, eager, lazy):
        self._d, self._lazy = dict(eager), dict(lazy)

    def __getitem__(self, k):
        if k not in self._d and k in self._lazy:
            self._d[k] = self._lazy.pop(k)()
        return self._d[k]

    def __iter__(self):
        return iter(list(self._d.keys()) + [k for k in self._lazy if k not in self._d])

    def __len__(self):
        return len(self._d) + len([k for k in self._lazy if k not in self._d])

    def packed_kwargs(self):
        """Only the packed entries (no padded tensor is built): ``model(**batch.packed_kwargs())``."""
        return {k: v for k, v in self._d.items()}


class DeviceCollate:
    def __init__(self, modality="atv", batch_first=True, speaker_onehot=False, n_speakers=2, device=None, pad_to=4):
        self.modalities = modality
        self.batch_first, self.speaker_onehot, self.n_speakers = batch_first, speaker_onehot, n_speakers
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.pad_to = pad_to
        self.copy_stream = None
        self.h2d_bytes = 0

    # ---- host side: one pass over the samples into pinned staging (the only CPU work)
    def pack_host(self, samples):
        lens = [len(s[0]["text"]) for s in samples]
        N = int(sum(lens))
        first = samples[0][0]
        dims = {"t": len(first["text"][0]), "a": len(first["audio"][0]), "v": len(first["visual"][0])}
        order = [m for m in self.modalities]
        D = sum(dims[m] for m in order)
        ld = (D + self.pad_to - 1) // self.pad_to * self.pad_to
        x = torch.zeros((N, ld), dtype=torch.float32).pin_memory()
        xn = x.numpy()
        spk = torch.empty(N, dtype=torch.int64).pin_memory()
        labels, emo, senti, sentences = [], [], [], []
        key = {"t": "text", "a": "audio", "v": "visual"}
        r = 0
        for (dic,), L in zip(samples, lens):
            c = 0
            for m in order:
                xn[r:r + L, c:c + dims[m]] = np.asarray(dic[key[m]], dtype=np.float32).reshape(L, dims[m])
                c += dims[m]
            spk[r:r + L] = torch.as_tensor(np.asarray(dic["speakers"])).argmax(dim=-1)        # mmbase.py:414
            labels.extend(dic["label"])
            if dic.get("emo_label") is not None:
                emo.append(np.asarray(dic["emo_label"]))
            if dic.get("senti2_label") is not None:
                senti.append(np.asarray(dic["senti2_label"]))
            if dic.get("sentence") is not None:
                sentences.append(dic["sentence"])
            r += L
        col_off, c = {}, 0
        for m in order:
            col_off[m] = (c, dims[m])
            c += dims[m]
        return dict(x=x, spk=spk, lengths=torch.tensor(lens, dtype=torch.int64), label=torch.tensor(labels).long().pin_memory(),
                    D=D, ld=ld, col_off=col_off, emo=emo, senti=senti, sentences=sentences)

    def __call__(self, samples):
        h = self.pack_host(samples)
        dev = self.device
        x_store = h["x"].to(dev, non_blocking=True)
        spk = h["spk"].to(dev, non_blocking=True)
        label = h["label"].to(dev, non_blocking=True)
        self.h2d_bytes += h["x"].numel() * 4 + h["spk"].numel() * 8 + h["label"].numel() * 8
        lengths = h["lengths"]
        B, Lmax, N, D = lengths.numel(), int(lengths.max()), x_store.size(0), h["D"]
        node_off = torch.zeros(B + 1, dtype=torch.int32)
        node_off[1:] = torch.cumsum(lengths, 0)
        node_off = node_off.to(dev, non_blocking=True)
        node_dlg = torch.repeat_interleave(torch.arange(B, dtype=torch.int32), lengths).to(dev, non_blocking=True)
        x_packed = x_store[:, :D]
        seq_first = not self.batch_first

        def unpack(c0, width):
            def make():
                shape = (Lmax, B, width) if seq_first else (B, Lmax, width)
                out = torch.zeros(shape, dtype=torch.float32, device=dev)
                src = x_store[:, c0:c0 + width]
                check(lib().ercg_unpack_rows(_p(src), x_store.stride(0), _p(node_off), _p(node_dlg), _p(out), width, Lmax, B,
                                             1 if seq_first else 0, N, width, _stream()), "ercg_unpack_rows")
                return out
            return make

        def masks():
            am = torch.empty((B, Lmax), dtype=torch.float32, device=dev)
            shape = (Lmax, B) if seq_first else (B, Lmax)
            ids = None if self.speaker_onehot else torch.empty(shape, dtype=torch.int64, device=dev)
            oh = torch.empty(shape + (self.n_speakers,), dtype=torch.float32, device=dev) if self.speaker_onehot else None
            check(lib().ercg_collate_masks(_p(node_off), _p(spk), B, Lmax, 1 if seq_first else 0,
                                           self.n_speakers if self.speaker_onehot else 0, _p(am), _p(ids), _p(oh), _stream()),
                  "ercg_collate_masks")
            return am, (oh if self.speaker_onehot else ids)

        cache = {}

        def mask_part(i):
            def make():
                if "m" not in cache:
                    cache["m"] = masks()
                return cache["m"][i]
            return make

        eager = {"text_length": lengths, "label": label, "x_packed": x_packed, "speaker_packed": spk}
        if h["sentences"]:
            eager["utterance_texts"] = h["sentences"]
        if h["emo"]:
            eager["emo_label"] = torch.from_numpy(np.concatenate([np.atleast_2d(e) for e in h["emo"]], 0)).to(dev)
        if h["senti"]:
            eager["senti2_label"] = torch.from_numpy(np.concatenate([np.atleast_1d(e) for e in h["senti"]], 0)).to(dev)
        lazy = {"attention_mask": mask_part(0), "speaker_tensor": mask_part(1), "input_tensor": unpack(0, D)}
        for m, name in (("t", "text_feature"), ("a", "audio_feature"), ("v", "visual_feature")):
            if m in h["col_off"]:
                lazy[name] = unpack(*h["col_off"][m])
            else:
                eager[name] = None
        return _LazyBatch(eager, lazy)
