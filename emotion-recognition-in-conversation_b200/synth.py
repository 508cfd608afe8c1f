"""Seeded synthetic inputs for the BASELINE.json configs.  CPU tensors; no reference data.

Shapes follow the reference's collate (track_mm/mmbase.py:354-455): ``input_tensor [B,Lmax,hidden_all]`` with
the modalities concatenated in the order of the letters of ``--modality`` (atv: a=100, t=768, v=512 for
iemocap-cogmen-sbert-4, mmbase.py:75-78,103-104), zero padding, ``speaker_tensor [B,Lmax]`` int64,
``text_length [B]`` int64, ``label [N]`` int64.
"""
import torch


def iemocap_lengths(B, gen):
    """L_d = clamp(round(49 + 21 z), 8, 110): IEMOCAP has 7433 utterances / 151 dialogues, cap 110."""
    z = torch.randn(B, generator=gen)
    return (49 + 21 * z).round().clamp(8, 110).to(torch.int64)


def mosei_lengths(total_utterances, gen, p=1.0 / 7.0, cap=40):
    """L_d = 1 + Geometric(p) clipped to cap (mean ~7 segments per video); draw until sum >= total."""
    est = int(total_utterances / 6.5) + 1024
    out, tot = [], 0
    while tot < total_utterances:
        u = torch.rand(est, generator=gen).clamp_(min=1e-12)
        L = (1 + torch.floor(torch.log(u) / torch.log(torch.tensor(1.0 - p)))).clamp(max=cap).to(torch.int64)
        out.append(L)
        tot += int(L.sum())
    L = torch.cat(out)
    cs = torch.cumsum(L, 0)
    n = int((cs < total_utterances).sum()) + 1
    return L[:n].contiguous()


def padded_batch(lengths, hidden_all, n_speakers, n_classes, gen, one_speaker=False):
    """Reference-layout batch (padded)."""
    B, Lmax = lengths.numel(), int(lengths.max())
    x = torch.randn(B, Lmax, hidden_all, generator=gen)
    spk = torch.zeros(B, Lmax, dtype=torch.int64) if one_speaker else torch.randint(0, n_speakers, (B, Lmax), generator=gen)
    mask = torch.arange(Lmax)[None, :] < lengths[:, None]
    x = x * mask[..., None]
    spk = spk * mask
    label = torch.randint(0, n_classes, (int(lengths.sum()),), generator=gen)
    return dict(input_tensor=x, speaker_tensor=spk, text_length=lengths, label=label, attention_mask=mask.float())


def packed_batch(lengths, hidden_all, n_speakers, n_classes, gen, one_speaker=False, ld=None):
    """Resident layout: utterance rows packed [N, hidden_all] (row stride ``ld`` >= hidden_all, default the next
    multiple of 4 so rows are 16-byte aligned), packed speakers [N]."""
    N = int(lengths.sum())
    ld = ld or (hidden_all + 3) // 4 * 4
    buf = torch.zeros(N, ld)
    buf[:, :hidden_all] = torch.randn(N, hidden_all, generator=gen)
    spk = torch.zeros(N, dtype=torch.int64) if one_speaker else torch.randint(0, n_speakers, (N,), generator=gen)
    label = torch.randint(0, n_classes, (N,), generator=gen)
    return dict(x_packed=buf[:, :hidden_all], x_storage=buf, speaker_packed=spk, text_length=lengths, label=label)


def config1(seed=0, B=32):
    """COGMEN, iemocap-cogmen-sbert-4 shape, modality atv (a=100,t=768,v=512), batch 32."""
    gen = torch.Generator().manual_seed(seed)
    return padded_batch(iemocap_lengths(B, gen), 1380, 2, 4, gen)


def config2(seed=0, B=32):
    """DialogueGCN 6-way IEMOCAP-shaped."""
    gen = torch.Generator().manual_seed(seed)
    return padded_batch(iemocap_lengths(B, gen), 1380, 2, 6, gen)


MOSEI_HIDDEN = 1443      # t=768 + a=640 + v=35  (mmbase.py:93,97-98,103-104)


def config5_lengths(total=1 << 20, seed=0):
    gen = torch.Generator().manual_seed(seed)
    return mosei_lengths(total, gen)


def to_bf16_rows(x, pitch_multiple=8):
    """[N, D] fp32 rows -> bf16 storage with a row pitch that is a multiple of 8 elements (16-byte aligned rows: what the
    bf16 input-feature kernels read through TMA).  Returns the [N, D] view of the padded buffer (round-to-nearest-even)."""
    N, D = x.shape
    ld = (D + pitch_multiple - 1) // pitch_multiple * pitch_multiple
    buf = torch.zeros((N, ld), dtype=torch.bfloat16, device=x.device)
    buf[:, :D] = x.to(torch.bfloat16)
    return buf[:, :D]
