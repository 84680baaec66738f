"""-m gpu: kernel-level parity through the C ABI — tcgen05 GEMM / implicit conv against fp32 torch math, and the
integer / boolean detector stages bit-exact against vectors the UNMODIFIED reference produced (tests/golden)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def eng_bare():
    from rgrg_b200 import Engine

    return Engine(0)


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("impl", [2, 0, 1, 3, 4])  # CUDA-core check, tcgen05 N tile 128 / 64 / 192 / 256
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (300, 200, 192), (928, 3072, 1024), (37, 800, 2048),
                                   (257, 50257, 1024), (1000, 64, 576)])
def test_gemm_matches_fp32_math(eng_bare, impl, M, N, K):
    A = _rand_bf16((M, K), 1)
    W = _rand_bf16((N, K), 2, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    out = eng_bare.gemm(A, W, bias, act=0, impl=impl)
    ref = A.float() @ W.float().T + bias
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


def test_gemm_fc6_shape_long_k(eng_bare):
    """fc6 of the box head: K = 131 072 (2048 k-blocks), the longest reduction on the path."""
    M, N, K = 64, 1024, 131072
    A = _rand_bf16((M, K), 11)
    W = _rand_bf16((N, K), 12, 0.003)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(13)).cuda()
    ref = A.float() @ W.float().T + bias
    for impl in (5, 4):
        out = eng_bare.gemm(A, W, bias, act=0, impl=impl)
        err = (out - ref).abs().max().item()
        assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "impl %d: max abs err %g" % (impl, err)
    out = eng_bare.gemm(A, W, bias, act=1, impl=5)
    assert (out - torch.relu(ref)).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(928, 1024, 1024), (928, 1024, 4096), (37, 1024, 4096), (300, 512, 256)])
def test_gemm_split_k_form(eng_bare, M, N, K):
    """The decoder's residual projections: 4 K slices -> fp32 partial sums -> reduce (c_proj K = 1024, mlp c_proj K = 4096)."""
    A = _rand_bf16((M, K), 21)
    W = _rand_bf16((N, K), 22, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(23)).cuda()
    out = eng_bare.gemm(A, W, bias, act=0, impl=6)
    ref = A.float() @ W.float().T + bias
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


@pytest.mark.parametrize("M,N,K", [(928, 4096, 1024), (928, 1024, 4096), (100, 256, 128), (300, 512, 320), (129, 1024, 1024), (1856, 4096, 1024)])
def test_gemm_cta_pair_kernel(eng_bare, M, N, K):
    """gemm_2cta.cuh (tcgen05.mma.cta_group::2, 256 x 256 pair-tiles, each CTA stages half of W): plain and split-K forms,
    odd numbers of M tiles (the pair's second CTA past the end), short K."""
    A = _rand_bf16((M, K), 31)
    W = _rand_bf16((N, K), 32, 0.05)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(33)).cuda()
    ref = A.float() @ W.float().T + bias
    tol = 2e-3 * max(1.0, ref.abs().max().item())
    out = eng_bare.gemm(A, W, bias, act=0, impl=7)
    assert (out - ref).abs().max().item() <= tol
    if (K // 64) % 4 == 0:
        out = eng_bare.gemm(A, W, bias, act=0, impl=8)
        assert (out - ref).abs().max().item() <= tol
    out = eng_bare.gemm(A, W, bias, act=2, impl=7)
    g = 0.5 * ref * (1 + torch.tanh(0.7978845608028654 * (ref + 0.044715 * ref ** 3)))
    assert (out - g).abs().max().item() <= 2e-3 * max(1.0, g.abs().max().item())
    # the TMA-store epilogue (default) and the register epilogue write the same bits
    try:
        eng_bare.set_option("epi_tma", 0)
        out0 = eng_bare.gemm(A, W, bias, act=2, impl=7)
        assert torch.equal(out, out0)
        if (K // 64) % 4 == 0:
            a1 = eng_bare.gemm(A, W, bias, act=0, impl=8)
            eng_bare.set_option("epi_tma", 1)
            assert torch.equal(a1, eng_bare.gemm(A, W, bias, act=0, impl=8))
    finally:
        eng_bare.set_option("epi_tma", 1)


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_epilogue_activations(eng_bare, act):
    A = _rand_bf16((200, 128), 4)
    W = _rand_bf16((256, 128), 5, 0.1)
    out = eng_bare.gemm(A, W, None, act=act, impl=0)
    ref = A.float() @ W.float().T
    ref = torch.relu(ref) if act == 1 else 0.5 * ref * (1 + torch.tanh(0.7978845608028654 * (ref + 0.044715 * ref ** 3)))
    assert (out - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 16, 64, 64), (1, 32, 128, 128), (2, 16, 2048, 256), (1, 128, 64, 64)])
def test_conv3x3_matches_torch(eng_bare, implicit, B, H, Cin, Cout):
    x = _rand_bf16((B, H, H, Cin), 6)
    w = _rand_bf16((Cout, 3, 3, Cin), 7, 0.05)  # tap-major K
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(8)).cuda()
    out = eng_bare.conv3x3(x, w.reshape(Cout, -1), bias, relu=True, implicit=implicit)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), "max abs err %g" % err


# ---- detector integer stages: need the engine's constant tables -> weights loaded once per module
@pytest.fixture(scope="module")
def eng(synth_sd):
    from rgrg_b200 import Engine

    e = Engine(0)
    e.load_state_dict(synth_sd)
    return e


def test_rpn_filter_bit_exact_on_reference_vectors(eng, golden):
    g = golden("rpn_filter.npz")
    obj = T(g["objectness"]).cuda()
    boxes, scores, count, topk, keep = eng.rpn_filter(obj, decoded=T(g["decoded"]).cuda())
    assert count.cpu().tolist() == list(g["count"])
    ref_topk = torch.topk(T(g["objectness"]), 1000, dim=1).indices
    assert torch.equal(topk.cpu().long(), ref_topk)
    for b in range(2):
        n = int(g["count"][b])
        assert torch.equal(boxes[b, :n].cpu(), T(g["boxes%d" % b]))
        assert torch.allclose(scores[b, :n].cpu(), T(g["scores%d" % b]), rtol=0, atol=1e-6)


def test_rpn_decode_on_device_matches_reference(eng, golden):
    g = golden("rpn_filter.npz")
    boxes, scores, count, topk, keep = eng.rpn_filter(T(g["objectness"]).cuda(), deltas=T(g["deltas"]).cuda())
    # device expf differs from the CPU's by <= 2 ulp: boxes agree to 1e-3 px, the keep list may differ only at IoU ties
    dec = T(g["decoded"])
    for b in range(2):
        n = int(count[b])
        sel = topk[b, keep[b, :n].long()].cpu().long()
        mine = boxes[b, :n].cpu()
        ref = dec[b, sel].clamp(0, 512)
        assert (mine - ref).abs().max().item() < 1e-3
        assert abs(n - int(g["count"][b])) <= 2


@pytest.mark.parametrize("separable", [1, 0])
def test_roi_align_matches_reference_kernel(eng, golden, separable):
    eng.set_option("roi_align_sep", separable)
    g = golden("roi_align.npz")
    feats = T(g["feats"]).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()  # NHWC
    rois = [T(g["rois0"]), T(g["rois1"])]
    boxes = torch.zeros(2, 1000, 4)
    for b in range(2):
        boxes[b, : rois[b].shape[0]] = rois[b]
    count = torch.tensor([r.shape[0] for r in rois], dtype=torch.int32)
    out = eng.roi_align(feats, boxes.cuda(), count.cuda())  # [P, 64 bins, C]
    import torchvision

    ref = torchvision.ops.roi_align(feats.float().permute(0, 3, 1, 2).cpu(), rois, (8, 8), 1.0 / 32, 2)  # bf16-rounded input
    ref = ref.permute(0, 2, 3, 1).reshape(ref.shape[0], 64, -1)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err < 0.02 * max(1.0, ref.abs().max().item())  # bf16 output rounding
    # and against the reference's fp32 output, looser (input + output rounding)
    ref32 = T(g["pooled"]).permute(0, 2, 3, 1).reshape(ref.shape[0], 64, -1)
    assert (out.float().cpu() - ref32).abs().max().item() < 0.03 * max(1.0, ref32.abs().max().item())
    eng.set_option("roi_align_sep", 1)


def test_roi_align_separable_edge_boxes(eng):
    """Boxes that leave the image, degenerate (zero-size) boxes, full-image boxes and sub-cell boxes: the separable kernel
    against torchvision's kernel on bf16-rounded inputs, 16x16 and 32x32 maps."""
    import torchvision

    for f, S in ((16, 512), (32, 1024)):
        g = torch.Generator().manual_seed(f)
        feats = torch.randn(2, f, f, 64, generator=g).to(torch.bfloat16)
        rois = []
        for b in range(2):
            r = torch.rand(40, 4, generator=g) * S
            x1, x2 = torch.minimum(r[:, 0], r[:, 2]), torch.maximum(r[:, 0], r[:, 2])
            y1, y2 = torch.minimum(r[:, 1], r[:, 3]), torch.maximum(r[:, 1], r[:, 3])
            bx = torch.stack([x1, y1, x2, y2], 1)
            bx[0] = torch.tensor([0.0, 0.0, float(S), float(S)])       # whole image
            bx[1] = torch.tensor([100.0, 100.0, 100.0, 100.0])         # zero size
            bx[2] = torch.tensor([S - 3.0, S - 3.0, float(S), float(S)])  # corner, sub-cell
            bx[3] = torch.tensor([0.0, 200.0, 5.0, 204.0])             # thin
            bx[4] = torch.tensor([10.0, 20.0, S - 1.0, 60.0])          # wide
            rois.append(bx)
        boxes = torch.zeros(2, 1000, 4)
        for b in range(2):
            boxes[b, :40] = rois[b]
        count = torch.tensor([40, 40], dtype=torch.int32)
        ref = torchvision.ops.roi_align(feats.float().permute(0, 3, 1, 2), rois, (8, 8), f / S, 2)
        ref = ref.permute(0, 2, 3, 1).reshape(80, 64, -1)
        for sep in (1, 0):
            eng.set_option("roi_align_sep", sep)
            out = eng.roi_align(feats.cuda(), boxes.cuda(), count.cuda(), image_size=S)
            err = (out.float().cpu() - ref).abs().max().item()
            assert err < 0.02 * max(1.0, ref.abs().max().item()), (f, sep, err)
    eng.set_option("roi_align_sep", 1)


def test_roi_tail_bit_exact_on_reference_vectors(eng, golden):
    g = golden("roi_tail.npz")
    props = [T(g["proposals0"]), T(g["proposals1"])]
    boxes = torch.zeros(2, 1000, 4)
    for b in range(2):
        boxes[b, : props[b].shape[0]] = props[b]
    count = torch.tensor([p.shape[0] for p in props], dtype=torch.int32)
    det, idx, scores, tb = eng.roi_tail(T(g["class_logits"]).cuda(), T(g["box_regression"]).cuda(), boxes.cuda(), count.cuda())
    assert torch.equal(det.cpu(), T(g["class_detected"]))
    assert torch.equal(idx.cpu().long(), T(g["top_idx"]))
    assert torch.allclose(scores.cpu(), T(g["top_scores"]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(tb.cpu(), T(g["top_region_boxes"]), rtol=0, atol=2e-3)


# ---- beam-search bookkeeping (K24-K26): device top-k / BeamSearchScorer.process / finalize against the reference loop
def _reference_beam_loop(logits_steps, sentences, nb, max_length, early):
    """language_model.py:545-607 with the model forward replaced by given logits; scorer = oracle/beam_scorer.py."""
    from beam_scorer import BeamSearchScorer

    V = logits_steps.shape[-1]
    scorer = BeamSearchScorer(batch_size=sentences, num_beams=nb, device=torch.device("cpu"), length_penalty=1.0,
                              do_early_stopping=early, num_beam_hyps_to_keep=1)
    ids = torch.full((sentences * nb, 1), 50256, dtype=torch.int64)
    beam_scores = torch.zeros(sentences, nb)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    cur_len = 1
    for t in range(logits_steps.shape[0]):
        scores = torch.log_softmax(logits_steps[t], dim=-1) + beam_scores[:, None]
        scores, tokens = torch.topk(scores.view(sentences, nb * V), 2 * nb, dim=1, largest=True, sorted=True)
        indices = torch.div(tokens, V, rounding_mode="floor")
        tokens = tokens % V
        out = scorer.process(ids, scores, tokens, indices, pad_token_id=50256, eos_token_id=50256)
        beam_scores = out["next_beam_scores"]
        ids = torch.cat([ids[out["next_beam_indices"], :], out["next_beam_tokens"].unsqueeze(-1)], dim=-1)
        cur_len += 1
        if scorer.is_done or cur_len >= max_length:
            break
    return scorer.finalize(ids, beam_scores, tokens, indices, pad_token_id=50256, eos_token_id=50256,
                           max_length=max_length)["sequences"]


@pytest.mark.parametrize("early", [True, False])
@pytest.mark.parametrize("eos_rate", [0.0, 0.15, 0.6])
def test_beam_bookkeeping_matches_reference_loop(eng_bare, early, eos_rate):
    sentences, nb, T, V = 5, 4, 9, 50257
    g = torch.Generator().manual_seed(int(eos_rate * 100) + int(early))
    logits = torch.randn(T - 1, sentences * nb, V, generator=g) * 2.0
    boost = torch.rand(T - 1, sentences * nb, generator=g) < eos_rate
    logits[:, :, 50256] = torch.where(boost, torch.full_like(logits[:, :, 0], 14.0), logits[:, :, 50256])
    ref = _reference_beam_loop(logits, sentences, nb, T, early)
    out = eng_bare.beam_bookkeeping(logits.cuda(), sentences, nb, T, early)
    assert out.shape == tuple(ref.shape), (out.shape, ref.shape)
    assert np.array_equal(out.astype(np.int64), ref.numpy())
