"""-m gpu: decoder parity at REAL cache lengths and on the EOS / early-stop path, through the C ABI.

* teacher-forced logits against tests/golden/lm_long.npz — 128 decode steps (cache length up to 129 keys) recorded from the
  UNMODIFIED reference: max |dlogit| <= 0.05 on the reference's top-8 tokens and on logsumexp at every step, arg-max
  equality wherever the reference's top-1 / top-2 margin exceeds 0.1 (SURVEY.md §8(d) tolerances);
* the same at 928 rows (BASELINE configs[1]: 32 images x 29 regions; 7 full M tiles + one 32-row tile);
* greedy / beam bookkeeping driven by crafted logits against ids the reference's own greedy_search / beam_search loops
  produced from the same logits (tests/golden/lm_crafted.npz): bit-exact ids and width;
* an EOS-emitting checkpoint variant: teacher-forced logits on the oracle's tokens (rows keep consuming pad tokens after
  they finish, language_model.py:636), free-running rows exact up to the first low-margin decision.
"""
import numpy as np
import pytest
import torch

import rgrg_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy
TOL = 0.05      # bf16 decoder vs fp32 reference (measured max 0.036 on logits of std 0.64)
MARGIN = 0.1


@pytest.fixture(scope="module")
def eng(synth_sd):
    from rgrg_b200 import Engine

    e = Engine(0)
    e.load_state_dict(synth_sd)
    return e


def _check_steps(logits, g, steps, rows=slice(None), col_rep=1):
    """logits [n, R, V] (cuda) against the golden record at `steps`; R may hold `col_rep` replicas of the golden rows."""
    worst = 0.0
    for t in steps:
        ref_idx = T(g["top_idx"][t]).long().cuda().repeat(col_rep, 1)
        ref_val = T(g["top_val"][t]).cuda().repeat(col_rep, 1)
        ref_lse = T(g["logsumexp"][t]).cuda().repeat(col_rep)
        l = logits[t]
        d = (l.gather(1, ref_idx) - ref_val).abs().max().item()
        d_lse = (torch.logsumexp(l, -1) - ref_lse).abs().max().item()
        worst = max(worst, d, d_lse)
        assert d <= TOL and d_lse <= TOL, "step %d: max |dlogit| %.4f, |dlogsumexp| %.4f" % (t, d, d_lse)
        confident = (ref_val[:, 0] - ref_val[:, 1]) > MARGIN
        assert torch.equal(l.argmax(-1)[confident], ref_idx[:, 0][confident]), "step %d: arg-max differs at a confident row" % t
    return worst


def test_logits_teacher_forced_at_real_cache_lengths(eng, golden):
    g = golden("lm_long.npz")
    feats, ids = T(g["feats"]).cuda(), T(g["ids"]).cuda()
    n = ids.shape[1] - 1  # 128 steps: cache length 2 .. 129, crossing the 16-key chunk boundary eight times
    logits = eng.lm_forced_logits(feats, ids[:, :n].contiguous())
    worst = _check_steps(logits, g, range(n))
    print("max deviation over %d steps: %.4f" % (n, worst))
    # the windows the round-1 review asked for explicitly
    for lo, hi in ((14, 18), (30, 34), (62, 64), (126, 128)):
        _check_steps(logits, g, range(lo, hi))


@pytest.mark.parametrize("opts", [dict(fused_attn=0), dict(attn_early=1), dict(attn_warps=8, attn_slots=4), dict(cuda_graph=0)])
def test_logits_teacher_forced_decode_variants(eng, golden, opts):
    g = golden("lm_long.npz")
    feats, ids = T(g["feats"]).cuda(), T(g["ids"]).cuda()
    try:
        for k, v in opts.items():
            eng.set_option(k, v)
        logits = eng.lm_forced_logits(feats, ids[:, :70].contiguous())  # cache length up to 71: five 16-key chunks
        _check_steps(logits, g, range(70))
    finally:
        for k, v in dict(fused_attn=1, attn_early=0, attn_warps=16, attn_slots=2, cuda_graph=1).items():
            eng.set_option(k, v)


def test_logits_teacher_forced_at_928_rows(eng, golden):
    g = golden("lm_long.npz")
    rep = 155  # 6 x 155 = 930 >= 928 rows
    feats = T(g["feats"]).repeat(rep, 1)[:928].contiguous().cuda()
    ids = T(g["ids"]).repeat(rep, 1)[:928, :20].contiguous().cuda()
    logits = eng.lm_forced_logits(feats, ids[:, :3].contiguous())  # [3, 928, V] = 560 MB
    for t in range(3):
        ref_idx = T(g["top_idx"][t]).long().repeat(rep, 1)[:928].cuda()
        ref_val = T(g["top_val"][t]).repeat(rep, 1)[:928].cuda()
        d = (logits[t].gather(1, ref_idx) - ref_val).abs().max().item()
        assert d <= TOL, "step %d: %.4f" % (t, d)
        # replicas of the same row in different M tiles produce identical logits
        assert torch.equal(logits[t][:6], logits[t][6 * 150: 6 * 151])


def test_greedy_bookkeeping_bit_exact_against_reference_loop(eng, golden):
    import crafted
    from rgrg_b200 import synth

    g = golden("lm_crafted.npz")
    for name, (seed, rows, max_length, kind) in crafted.GREEDY_CASES.items():
        mask = crafted.eos_schedule(kind, max_length - 1, rows)
        logits = synth.crafted_logits(seed, max_length - 1, rows, mask).cuda()
        out = eng.greedy_bookkeeping(logits, max_length)
        ref = g["greedy_%s_ids" % name]
        assert out.shape == ref.shape, (name, out.shape, ref.shape)
        assert np.array_equal(out, ref), name


def test_beam_bookkeeping_adversarial_cases_against_reference_loop(eng, golden):
    import crafted

    g = golden("lm_crafted.npz")
    for name, (seed, sentences, nb, max_length, es, kind) in crafted.BEAM_CASES.items():
        logits = crafted.beam_crafted_logits(seed, sentences, nb, max_length, kind)
        out = eng.beam_bookkeeping(logits.cuda(), sentences, nb, max_length, es)
        if kind == "ties":
            # exact ties: torch.topk's order of equal candidates is unspecified (the golden pins torch's CPU kernel, which the
            # CPU oracle test checks); the engine's rule is "lowest flat index first" = the reference loop with a stable sort
            ref = O.beam_search({}, torch.zeros(sentences, 1024), max_length, nb, es, given_logits=logits, stable_ties=True).numpy()
        else:
            ref = g["beam_%s_ids" % name]
        assert out.shape == ref.shape, (name, out.shape, ref.shape)
        assert np.array_equal(out, ref), name


def test_eos_checkpoint_logits_and_early_stop(synth_sd, golden):
    """A checkpoint whose final LayerNorm bias is shifted along wte[EOS], so rows emit EOS at staggered steps."""
    from rgrg_b200 import Engine

    sd = dict(synth_sd)
    w = sd["language_model.wte.weight"][50256]
    sd["language_model.final_layernorm.bias"] = sd["language_model.final_layernorm.bias"] + 2.5 * w / (w @ w)
    feats = T(golden("selection.npz")["selected_features"])[:12].contiguous()
    rec = {}
    with torch.no_grad():
        ref = O.lm_generate(sd, feats, max_length=40, record=rec)
    finished_at = [(ref[r, 1:] == 50256).nonzero()[0].item() + 1 if (ref[r, 1:] == 50256).any() else None for r in range(12)]
    assert sum(f is not None for f in finished_at) >= 6 and len(set(finished_at)) >= 4, finished_at  # staggered finishes
    e = Engine(0)
    e.load_state_dict(sd)
    n = ref.shape[1] - 1
    logits = e.lm_forced_logits(feats.cuda(), ref[:, :n].to(torch.int32).cuda()).cpu()
    for t in range(n):
        d = (logits[t] - rec["logits"][t]).abs().max().item()
        assert d <= TOL, "step %d: %.4f" % (t, d)
        top2 = rec["logits"][t].topk(2, -1).values
        confident = (top2[:, 0] - top2[:, 1]) > MARGIN
        assert torch.equal(logits[t].argmax(-1)[confident], rec["logits"][t].argmax(-1)[confident])
    # free-running: a row must equal the oracle's row up to its first low-margin decision; rows whose every decision is
    # confident must match completely, EOS padding included
    out = e.lm_generate(feats.cuda(), 40)
    margins = torch.stack([l.topk(2, -1).values[:, 0] - l.topk(2, -1).values[:, 1] for l in rec["logits"]])  # [steps, rows]
    all_exact = True
    for r in range(12):
        last = finished_at[r] if finished_at[r] is not None else n
        low = (margins[:last, r] < MARGIN).nonzero()
        safe = int(low[0]) if len(low) else last
        w = min(out.shape[1], ref.shape[1], safe + 1)
        assert np.array_equal(out[r, :w], ref[r, :w].numpy()), "row %d differs before its first low-margin step %d" % (r, safe)
        if safe == last:
            assert np.array_equal(out[r, :min(out.shape[1], ref.shape[1])], ref[r, :min(out.shape[1], ref.shape[1])].numpy())
        else:
            all_exact = False
    if all_exact:
        assert out.shape == tuple(ref.shape)
    e.close()
