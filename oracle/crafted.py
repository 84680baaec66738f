"""Crafted logits for the search-loop bookkeeping fixtures (TEST INFRASTRUCTURE ONLY).

oracle/make_golden.py feeds them to the UNMODIFIED reference's greedy_search / beam_search (the model forward replaced by
these logits, nothing else) and commits the resulting ids in tests/golden/lm_crafted.npz; the tests rebuild the same
logits from the seeds and drive the CPU oracle and the CUDA bookkeeping kernels with them."""
import torch

GREEDY_CASES = {
    # name: (seed, rows, max_length, EOS schedule).  Schedules cover: staggered finishes with an early exit, rows that never
    # finish (width = max_length), everybody finishing at the same step, EOS on the very first step, and finishes that
    # straddle the engine's every-8-steps host check (steps 7, 8, 9) as well as the last possible step.
    "staggered": (11, 12, 24, "staggered"),
    "never_all": (12, 7, 12, "some_never"),
    "same_step": (13, 5, 20, "same_step"),
    "first_step": (14, 4, 9, "first_step"),
    "around_8": (15, 6, 30, "around_8"),
    "last_step": (16, 5, 10, "last_step"),
}


def eos_schedule(kind, steps, rows):
    m = torch.zeros(steps, rows, dtype=torch.bool)
    if kind == "staggered":
        for r in range(rows):
            m[2 + (5 * r) % 13, r] = True
            m[min(steps - 1, 3 + (5 * r) % 13 + 4), r] = True  # a second EOS after the row is finished: must stay padded
    elif kind == "some_never":
        for r in range(0, rows, 2):
            m[1 + r, r] = True
    elif kind == "same_step":
        m[6, :] = True
    elif kind == "first_step":
        m[0, :] = True
    elif kind == "around_8":
        for r in range(rows):
            m[6 + r % 3, r] = True  # last finishes at step index 6, 7, 8
    elif kind == "last_step":
        m[steps - 1, :] = True
        m[3, 0] = True
    return m


BEAM_CASES = {
    # name: (seed, sentences, beams, max_length, early_stopping, kind)
    "ties_es1": (21, 3, 4, 8, True, "ties"),
    "ties_es0": (22, 3, 4, 8, False, "ties"),
    "eos_low_rank_es1": (23, 4, 4, 9, True, "eos_low_rank"),
    "eos_low_rank_es0": (24, 4, 4, 9, False, "eos_low_rank"),
    "all_finish_es1": (25, 3, 4, 8, True, "all_finish"),
    "all_finish_es0": (26, 3, 4, 8, False, "all_finish"),
    "eos_first_es1": (27, 2, 4, 6, True, "eos_first"),
}


def beam_crafted_logits(seed, sentences, nb, max_length, kind):
    """Adversarial inputs for the beam bookkeeping (language_model.py:556-607): exact score ties across beams, EOS
    candidates ranked at or beyond num_beams (must be skipped), every beam finishing at the same step."""
    from rgrg_b200 import synth

    steps, rows, V = max_length - 1, sentences * nb, 50257
    logits = synth.crafted_logits(seed, steps, rows, None)
    if kind == "ties":
        # step 0: the two best tokens of beam 0 tie exactly; later steps: every beam of a sentence sees the same logits,
        # so candidates of equal-score beams tie across beams
        for t in range(steps):
            for s_ in range(sentences):
                logits[t, s_ * nb:(s_ + 1) * nb] = logits[t, s_ * nb]
        top = logits[0].topk(3, dim=-1)
        for r in range(rows):
            logits[0, r, top.indices[r, 1]] = top.values[r, 0]
            logits[0, r, top.indices[r, 2]] = top.values[r, 0]
    elif kind == "eos_low_rank":
        # EOS gets a logit that usually ranks it between num_beams and 2*num_beams among the sentence's candidates
        for t in range(1, steps):
            for r in range(rows):
                kth = logits[t, r].topk(3).values[2 - (r + t) % 3]
                logits[t, r, V - 1] = kth - 0.01 * (1 + (r + t) % 2)  # never an exact tie
    elif kind == "all_finish":
        logits[3, :, V - 1] = 30.0
    elif kind == "eos_first":
        logits[0, :, V - 1] = 30.0
    return logits
