"""n4 — report assembly (row -> image map, exact / soft duplicate removal): the product's implementation against the
control-flow restatement of evaluate_language_model.py:985-1091 in oracle/report_oracle.py, on randomised inputs with
stub sentence-splitter / similarity assets (the real ones need spaCy and DistilBERT files)."""
import random
import zlib

import numpy as np

import report_oracle as RO
from rgrg_b200 import report_assembly as RA


class _Span:
    def __init__(self, text):
        self.text = text


class _Doc:
    def __init__(self, sents):
        self.sents = [_Span(s) for s in sents]


def _splitter(text):  # spaCy-like: doc.sents -> spans with .text; splits after every full stop
    parts = [p.strip() for p in text.replace(".", ".|").split("|")]
    return _Doc([p for p in parts if p])


class _Score:  # evaluate-like: compute(...)["f1"][0]; deterministic pseudo-similarity
    def compute(self, lang, predictions, references, model_type):
        a, b = predictions[0], references[0]
        h = zlib.crc32((min(a, b) + "#" + max(a, b)).encode()) % 1000
        return {"f1": [h / 1000.0]}


VOCAB = ["The heart is normal.", "No pleural effusion.", "Lungs are clear.", "Mild cardiomegaly.", "No pneumothorax.",
         "The heart is unremarkable.", "Stable appearance.", "Lines and tubes in place."]


def _case(seed, B):
    rng = random.Random(seed)
    sel = np.array([[rng.random() < 0.5 for _ in range(29)] for _ in range(B)])
    sents = []
    for b in range(B):
        for _ in range(int(sel[b].sum())):
            n = rng.choice([1, 1, 1, 2])
            sents.append(" ".join(rng.choice(VOCAB) for _ in range(n)))
    return sel, sents


def test_row_image_map():
    sel = np.zeros((3, 29), dtype=bool)
    sel[0, [1, 5]] = True
    sel[2, [0, 7, 28]] = True
    img, reg, off = RA.row_image_map(sel)
    assert img.tolist() == [0, 0, 2, 2, 2] and reg.tolist() == [1, 5, 0, 7, 28] and off.tolist() == [0, 2, 2, 5]


def test_reports_match_reference_control_flow():
    for seed in range(40):
        sel, sents = _case(seed, B=1 + seed % 4)
        for thr in (0.3, 0.6, 0.9):
            ref = RO.get_generated_reports(sents, sel, _splitter, thr, _Score())
            out = RA.get_generated_reports(sents, sel, _splitter, thr, _Score())
            assert out[0] == ref[0], (seed, thr)
            assert [dict(d) for d in out[1]] == [dict(d) for d in ref[1]], (seed, thr)


def test_without_assets_only_exact_duplicates_go():
    sel = np.zeros((1, 29), dtype=bool)
    sel[0, :4] = True
    reports, removed = RA.get_generated_reports(["a b.", "c.", "a b.", "d."], sel)
    assert reports == ["a b. c. d."] and removed == [{}]


def test_dedup_rows_on_token_ids():
    sel = np.zeros((2, 29), dtype=bool)
    sel[0, :3] = True
    sel[1, :2] = True
    E = RA.EOS
    ids = np.array([[E, 5, 6, E, E], [E, 7, E, E, E], [E, 5, 6, E, 9],   # row 2 repeats row 0 (tokens after EOS do not count)
                    [E, 5, 6, E, E], [E, 5, 6, 7, E]])                   # other image: not a duplicate of image 0
    assert RA.dedup_rows(ids, sel).tolist() == [True, True, False, True, True]
    reports, _ = RA.reports_from_ids(ids, sel, decode=lambda rows: [" ".join("t%d" % t for t in r if t != E) + "." for r in rows])
    assert reports == ["t5 t6. t7.", "t5 t6. t5 t6 t7."]
