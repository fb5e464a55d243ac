"""Analogy golden vectors from the UNMODIFIED reference (torchkge v0.17.7 at /root/reference,
models/bilinear.py:559-763).  Run in the authoring container only:

    PYTHONPATH=/root/reference python tests/golden/make_golden_analogy.py

Writes toy_analogy.npz / syn_analogy.npz in the format of make_golden.py (link-prediction ranks,
dense inference scores, scoring_function, Model.forward + MarginLoss + gradients) and
rel_toy_analogy.npz / rel_syn_analogy.npz in the format of make_golden_rel.py (relation-prediction
ranks, dense relation scores), by running those two generators with an Analogy constructor.
``dim`` in the files is emb_dim (scalar_dim = complex_dim = emb_dim / 2, the default scalar_share).
"""
import os
import sys

import pandas as pd

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from torchkge.models import AnalogyModel  # noqa: E402

import make_golden  # noqa: E402
import make_golden_rel  # noqa: E402


def build(kind, d, n_ent, n_rel):
    assert kind == "analogy"
    return AnalogyModel(d, n_ent, n_rel)


def main():
    make_golden.build = build
    make_golden_rel.build = build
    toy = pd.DataFrame(make_golden.TOY, columns=["from", "to", "rel"])
    make_golden.run_case("toy_analogy", "analogy", 100, toy, 9, seed=7, b_size=9)
    df = make_golden.synthetic_df(300, 9, 2500, seed=11)
    make_golden.run_case("syn_analogy", "analogy", 72, df, 160, seed=13, b_size=40)
    make_golden_rel.CASES = ["toy_analogy", "syn_analogy"]
    make_golden_rel.main()


if __name__ == "__main__":
    main()
