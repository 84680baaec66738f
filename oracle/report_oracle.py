"""CPU oracle of the report-assembly step after the path (TEST INFRASTRUCTURE ONLY).

Restates `get_generated_reports` (src/full_model/evaluate_full_model/evaluate_language_model.py:985-1091) with its
nested `remove_duplicate_generated_sentences` (:1006-1064; the same function appears at
src/full_model/generate_reports_for_images.py:42-97) control flow for control flow, so that the product's differently
structured implementation (rgrg_b200/report_assembly.py) can be compared with it on arbitrary inputs.  The reference
module itself cannot be imported here: it pulls spaCy / evaluate / pycocoevalcap at import time."""
from collections import defaultdict

import numpy as np


def remove_duplicate_generated_sentences(gen_report_single_image, bert_score, sentence_tokenizer, bertscore_threshold):
    def in_removed(gen_sent, table):  # :1007-1012
        for lst in table.values():
            if gen_sent in lst:
                return True
        return False

    sents = [s.text for s in sentence_tokenizer(gen_report_single_image).sents]  # :1017-1020
    sents = list(dict.fromkeys(sents))  # :1024
    table = defaultdict(list)  # :1031
    for i in range(len(sents)):  # :1039-1059
        s1 = sents[i]
        for j in range(i + 1, len(sents)):
            if in_removed(s1, table):
                break
            s2 = sents[j]
            if in_removed(s2, table):
                continue
            res = bert_score.compute(lang="en", predictions=[s1], references=[s2], model_type="distilbert-base-uncased")
            if res["f1"][0] > bertscore_threshold:
                if len(s1) > len(s2):
                    table[s1].append(s2)
                else:
                    table[s2].append(s1)
    report = " ".join(s for s in sents if not in_removed(s, table))  # :1061-1063
    return report, table


def get_generated_reports(generated_sentences_for_selected_regions, selected_regions, sentence_tokenizer, bertscore_threshold,
                          bert_score):
    reports, removed = [], []
    curr = 0
    for sel_single in selected_regions:  # :1070-1089
        n = int(np.sum(sel_single))
        sents = generated_sentences_for_selected_regions[curr:curr + n]
        curr += n
        text = " ".join(s for s in sents)
        text, table = remove_duplicate_generated_sentences(text, bert_score, sentence_tokenizer, bertscore_threshold)
        reports.append(text)
        removed.append(table)
    return reports, removed
