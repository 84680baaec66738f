"""Report assembly after the path (n4): token rows -> per-image reports.

Host-side index work that follows `ReportGenerationModel.generate()` in the reference's callers
(`get_generated_reports`, src/full_model/evaluate_full_model/evaluate_language_model.py:985-1091;
`convert_generated_sentences_to_report`, src/full_model/generate_reports_for_images.py:42-104):
rows of the decoder output belong to images in row-major order of `selected_regions` (image-major, region-minor), a
report is the image's sentences joined by spaces after exact-duplicate removal (an insertion-ordered set) and an
optional soft de-duplication driven by a pairwise similarity score (BERTScore F1 > threshold: drop the shorter one).

The tokenizer / sentence splitter / BERTScore model are assets of the caller (they need files that are not part of
this engine); they are passed in as callables with the interfaces the reference uses.  Without them the functions fall
back to "one generated string = one sentence" and to exact-duplicate removal only.  `dedup_rows` does the exact-duplicate
step on TOKEN IDS before any text exists, so duplicate rows never reach the tokenizer.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

EOS = 50256


def row_image_map(selected_regions) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """selected_regions bool [B, 29] -> (image index [R], region index [R], first row of every image [B + 1]) for the R
    decoder rows (`selected_region_features = top_region_features[selected_regions]`,
    binary_classifier_region_selection.py:61: image-major, region-minor)."""
    sel = np.asarray(selected_regions, dtype=bool)
    img, reg = np.nonzero(sel)
    offsets = np.zeros(sel.shape[0] + 1, dtype=np.int64)
    np.cumsum(sel.sum(axis=1), out=offsets[1:])
    return img.astype(np.int64), reg.astype(np.int64), offsets


def dedup_rows(output_ids, selected_regions) -> np.ndarray:
    """Exact-duplicate filter on token ids: keep[r] is False when an EARLIER row of the same image generated the same token
    sequence (compared up to the first EOS after BOS).  Equal ids decode to equal strings, so this removes the rows the
    reference's `list(dict.fromkeys(sentences))` would drop, before they are decoded."""
    ids = np.asarray(output_ids)
    img, _, _ = row_image_map(selected_regions)
    if ids.shape[0] != img.shape[0]:
        raise ValueError("output_ids has %d rows but selected_regions selects %d" % (ids.shape[0], img.shape[0]))
    keep = np.ones(ids.shape[0], dtype=bool)
    seen = set()
    for r in range(ids.shape[0]):
        row = ids[r, 1:]
        end = np.nonzero(row == EOS)[0]
        key = (int(img[r]), row[: int(end[0])].tobytes() if len(end) else row.tobytes())
        if key in seen:
            keep[r] = False
        else:
            seen.add(key)
    return keep


def _split_sentences(text: str, sentence_tokenizer) -> List[str]:
    if sentence_tokenizer is None:
        return [text] if text else []
    return [s.text for s in sentence_tokenizer(text).sents]


def remove_duplicate_sentences(sentences: Sequence[str], similarity: Optional[Callable[[str, str], float]] = None,
                               threshold: float = 0.9) -> Tuple[List[str], Dict[str, List[str]]]:
    """Exact duplicates first (insertion-ordered set), then the reference's pairwise soft de-duplication: walking the
    pairs (i < j) in order, a pair whose similarity exceeds `threshold` loses its SHORTER sentence (the second one on
    equal length keeps, the first goes); a sentence that has been removed neither starts nor joins further comparisons.
    Returns the kept sentences and {kept sentence: [sentences removed in its favour]}."""
    uniq = list(dict.fromkeys(sentences))
    removed_for: Dict[str, List[str]] = defaultdict(list)
    gone = set()
    if similarity is not None:
        for i, a in enumerate(uniq):
            for b in uniq[i + 1:]:
                if a in gone:
                    break
                if b in gone:
                    continue
                if similarity(a, b) > threshold:
                    keep, drop = (a, b) if len(a) > len(b) else (b, a)
                    removed_for[keep].append(drop)
                    gone.add(drop)
    return [s for s in uniq if s not in gone], removed_for


def get_generated_reports(generated_sentences_for_selected_regions: Sequence[str], selected_regions, sentence_tokenizer=None,
                          bertscore_threshold: float = 0.9, bert_score=None):
    """Drop-in for evaluate_language_model.py:985-1091 (same arguments, same return pair): list of B reports and list of
    B dicts {kept sentence: [removed similar sentences]}.  `bert_score` follows the `evaluate` module interface
    (`compute(lang, predictions, references, model_type)["f1"][0]`); None disables the soft de-duplication."""
    _, _, offsets = row_image_map(selected_regions)
    if offsets[-1] != len(generated_sentences_for_selected_regions):
        raise ValueError("%d sentences for %d selected regions" % (len(generated_sentences_for_selected_regions), offsets[-1]))
    similarity = None
    if bert_score is not None:
        def similarity(a, b):
            return bert_score.compute(lang="en", predictions=[a], references=[b], model_type="distilbert-base-uncased")["f1"][0]
    reports, removed = [], []
    for b in range(len(offsets) - 1):
        mine = list(generated_sentences_for_selected_regions[offsets[b]:offsets[b + 1]])
        # the reference joins the image's strings and re-splits them with spaCy; without a splitter every region's string
        # counts as one sentence
        sents = mine if sentence_tokenizer is None else _split_sentences(" ".join(mine), sentence_tokenizer)
        kept, dropped = remove_duplicate_sentences(sents, similarity, bertscore_threshold)
        reports.append(" ".join(kept))
        removed.append(dropped)
    return reports, removed


def reports_from_ids(output_ids, selected_regions, decode: Callable[[Sequence[Sequence[int]]], List[str]], sentence_tokenizer=None,
                     bertscore_threshold: float = 0.9, bert_score=None):
    """generate() output -> reports: duplicate rows are dropped on token ids, the rest goes through `decode`
    (`tokenizer.batch_decode(..., skip_special_tokens=True)` in the reference, generate_reports_for_images.py:118) and
    `get_generated_reports`."""
    ids = np.asarray(output_ids)
    sel = np.asarray(selected_regions, dtype=bool)
    keep = dedup_rows(ids, sel)
    img, reg, _ = row_image_map(sel)
    kept_sel = np.zeros_like(sel)
    kept_sel[img[keep], reg[keep]] = True
    texts = decode([row.tolist() for row in ids[keep]])
    return get_generated_reports(texts, kept_sel, sentence_tokenizer, bertscore_threshold, bert_score)
