"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle."""
import numpy as np
import torch


def run():
    import rgrg_oracle as O  # oracle/ is on sys.path (set by __graft_entry__.smoke): used as the checker only

    from . import ReportGenerationModel, synth

    sd = synth.make_state_dict(0)
    images = synth.synthetic_images(1, 512, seed=1001)
    model = ReportGenerationModel(pretrain_without_lm_model=True)
    model.load_state_dict(sd)
    model.to(torch.device("cuda", 0))
    model.eval()
    out = model.generate(images, max_length=4)
    assert out != -1
    ids, selected, detections, class_detected = out
    with torch.no_grad():
        det = O.detect(sd, images)
        sel, feats, _ = O.region_selection(sd, det["top_region_features"], det["class_detected"])
    assert torch.equal(class_detected.cpu(), det["class_detected"]), "class_detected differs from the oracle"
    assert torch.equal(selected.cpu(), sel), "selected regions differ from the oracle"
    assert ids.shape == (int(sel.sum()), 4)
    rec = {}
    ref_ids = O.lm_generate(sd, feats[:4].contiguous(), max_length=3, record=rec)
    logits = model._engine().lm_forced_logits(feats[:4].cuda(), ref_ids[:, :-1].to(torch.int32).cuda()).cpu()
    err = max((logits[t] - rec["logits"][t]).abs().max().item() for t in range(len(rec["logits"])))
    assert err <= 0.05, "decoder logits off by %g" % err
    print("smoke ok: R=%d regions, %d launches, max |dlogit| %.4f" % (ids.shape[0], model._engine().kernel_launches, err))
