#!/usr/bin/env python
"""python train_mm.py --module=cogmen --dataset=iemocap-cogmen-sbert-4 --modality=atv --device=0   (reference: train_mm.py:16-25)
Launcher shim: the reference's entry point on the libercgraph modules, without lumo / accelerate / mmdatasets."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import erc_b200  # noqa: E402,F401
from erc_b200.train_mm import main  # noqa: E402

if __name__ == "__main__":
    main()
