"""CPU restatement of the reference collate ``ERCCollate.__call__`` (track_mm/mmbase.py:354-455) and a synthetic sample
generator in the data sets' sample format.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); pinned against the real
class in tests/test_collate.py."""
import numpy as np
import torch


def synthetic_samples(lengths, dims, n_speakers, n_classes, seed=0, with_sentence=False):
    """Samples as the reference data sets hand them to the collate: ``[(dic,), ...]`` with per-utterance lists of float32
    vectors under 'text' / 'audio' / 'visual', one-hot 'speakers' lists and a 'label' list (mmbase.py:374-384)."""
    rng = np.random.default_rng(seed)
    dt, da, dv = dims
    out = []
    for L in lengths:
        spk = rng.integers(0, n_speakers, size=L)
        dic = {"text": [rng.standard_normal(dt).astype(np.float32) for _ in range(L)],
               "audio": [rng.standard_normal(da).astype(np.float32) for _ in range(L)],
               "visual": [rng.standard_normal(dv).astype(np.float32) for _ in range(L)],
               "speakers": [[1 if s == k else 0 for k in range(n_speakers)] for s in spk],
               "label": [int(v) for v in rng.integers(0, n_classes, size=L)]}
        if with_sentence:
            dic["sentence"] = ["utt %d" % i for i in range(L)]
        out.append((dic,))
    return out


def collate(samples, modality="atv", batch_first=True, speaker_onehot=False, n_speakers=2):
    """Same keys / shapes / dtypes as the reference's dict."""
    lens = np.array([len(s[0]["text"]) for s in samples], dtype=np.int64)
    B, mx = len(samples), int(lens.max())
    key = {"t": "text", "a": "audio", "v": "visual"}

    def padded(name_or_list):
        rows = []
        for (dic,), L in zip(samples, lens):
            if isinstance(name_or_list, str):
                m = np.stack([np.asarray(v, dtype=np.float32) for v in dic[name_or_list]])
            else:
                m = np.concatenate([np.stack([np.asarray(v, dtype=np.float32) for v in dic[key[c]]]) for c in name_or_list], 1)
            rows.append(np.concatenate([m, np.zeros((mx - L, m.shape[1]), dtype=np.float32)], 0))
        return torch.from_numpy(np.stack(rows, 0 if batch_first else 1))

    am = torch.from_numpy((np.arange(mx)[None, :] < lens[:, None]).astype(np.float32))
    spk = np.zeros((B, mx), dtype=np.int64)
    for i, ((dic,), L) in enumerate(zip(samples, lens)):
        spk[i, :L] = np.asarray(dic["speakers"]).argmax(-1)
    spk = torch.from_numpy(spk)
    if not batch_first:
        spk = spk.transpose(0, 1)
    if speaker_onehot:
        spk = torch.zeros(*spk.shape, n_speakers).scatter_(-1, spk.unsqueeze(-1), 1)
    data = {"attention_mask": am, "text_length": torch.from_numpy(lens),
            "text_feature": padded("text") if "t" in modality else None,
            "audio_feature": padded("audio") if "a" in modality else None,
            "visual_feature": padded("visual") if "v" in modality else None,
            "input_tensor": padded(list(modality)), "speaker_tensor": spk,
            "label": torch.tensor([v for (dic,) in samples for v in dic["label"]]).long()}
    if samples[0][0].get("sentence") is not None:
        data["utterance_texts"] = [dic["sentence"] for (dic,) in samples]
    return data
