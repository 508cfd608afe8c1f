"""the ``train_mm.py --module=... --modality=atv`` entry point (reference train_mm.py:16-25,
mmbase.py:483-499): argument parsing and the data-set -> feature-size rules on the CPU; a short synthetic run of every module
on the GPU (loss falls, test pass with sklearn metrics, checkpoint with the reference's state_dict keys)."""
import os

import pytest
import torch


def test_params_follow_the_reference_rules():
    import erc_b200  # noqa: F401
    from erc_b200.train_mm import parse_args, resolve_params
    a = parse_args(["--module=cogmen", "--dataset=iemocap-cogmen-sbert-4", "--modality=atv", "--device=0", "--train.batch_size", "8", "--debug"])
    assert a == {"module": "cogmen", "dataset": "iemocap-cogmen-sbert-4", "modality": "atv", "device": 0, "train.batch_size": 8, "debug": True}
    p = resolve_params(a)
    assert (p["hidden_audio"], p["hidden_text"], p["hidden_visual"], p["hidden_all"]) == (100, 768, 512, 1380)     # mmbase.py:75-78,103-104
    assert p["n_classes"] == 4 and p["batch_size"] == 8 and p["lr"] == 1e-4 and p["weight_decay"] == 1e-8 and p["epoch"] == 55
    p = resolve_params(parse_args(["--module=cogmen", "--dataset=mosei-emo-sbert-fbank-6"]))
    assert (p["hidden_text"], p["hidden_audio"], p["hidden_visual"], p["hidden_all"]) == (768, 640, 35, 1443)
    p = resolve_params(parse_args(["--module=mmgcn", "--dataset=meld-mmgcn-7"]))
    assert p["n_speakers"] == 9 and (p["hidden_audio"], p["hidden_text"], p["hidden_visual"]) == (300, 600, 342)
    assert p["batch_first"] is False and p["speaker_onehot"] is True and p["batch_size"] == 16
    p = resolve_params(parse_args(["--module=dagerc", "--dataset=iemocap-cogmen-6", "--reimplement"]))
    assert p["optim"] == "AdamW" and p["lr"] == 5e-4 and p["dropout"] == 0.2 and p["batch_size"] == 16 and p["speaker_onehot"]
    with pytest.raises(SystemExit):
        resolve_params(parse_args(["--module=nope"]))


@pytest.mark.gpu
@pytest.mark.parametrize("module,extra", [("cogmen", ["--dataset=iemocap-cogmen-sbert-4"]), ("dgcn", ["--dataset=iemocap-cogmen-6"]), ("dgcnv2", ["--dataset=iemocap-cogmen-6", "--max_len=40"]),
                                          ("mmgcn", ["--dataset=iemocap-cogmen-6"]), ("dagerc", ["--dataset=iemocap-cogmen-6", "--max_len=24"])])
def test_synthetic_epochs_run_end_to_end(module, extra, tmp_path):
    import erc_b200  # noqa: F401
    from erc_b200.train_mm import main, build_model
    lines = []
    out = main(["--module=" + module, "--modality=atv", "--epoch=3", "--train_dialogues=24", "--test_dialogues=8",
                "--train.batch_size=8", "--save_dir=" + str(tmp_path), "--lr=0.001"] + extra, log=lambda *a: lines.append(" ".join(str(x) for x in a)))
    h = out["history"]
    assert len(h) == 3 and all(torch.isfinite(torch.tensor(r["train_Lall"])) for r in h)
    assert h[-1]["train_Lall"] < h[0]["train_Lall"]
    assert 0.0 <= out["best"]["acc"] <= 1.0 and len(out["best"]["cm"]) == out["params"]["n_classes"]
    for name in ("best_model.ckpt", "last_model.ckpt"):
        sd = torch.load(os.path.join(str(tmp_path), name))
        fresh = build_model(out["params"], torch.device("cuda"))
        fresh.load_state_dict(sd, strict=True)
    assert any(l.startswith("Best Results") for l in lines)
