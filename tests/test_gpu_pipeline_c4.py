"""BASELINE config C4 on the GPU: classifier -> thresholds -> routed CDAN weight sets on a mixed synthetic batch.  The routed
result must equal enhancing every image on its own through the flagged weight sets, bit for bit (the forward is
batch-independent), and the classifier on the GPU must agree with the CPU module (<= 1e-3 on the logits: cuDNN vs CPU)."""
import pytest
import torch

from oracle.stress_init import stress_state_dict

pytestmark = pytest.mark.gpu


def test_routed_pipeline_equals_per_image_enhancement(cuda_device):
    from classification.multilabel_classifier import DEGRADATIONS, MultiHeadClassifier, predict_probs
    from models.cdan import CDAN
    from routing import ENHANCER_CLASSES, MultiDegradationPipeline, synthetic_mixed_batch
    torch.manual_seed(11)
    clf_cpu = MultiHeadClassifier().eval()
    clf = MultiHeadClassifier().eval()
    clf.load_state_dict(clf_cpu.state_dict())
    clf = clf.to(cuda_device)
    enh = {}
    for k, name in enumerate(ENHANCER_CLASSES):
        net = CDAN().set_compute_dtype("bf16")
        net.load_state_dict(stress_state_dict(300 + k))
        enh[name] = net.to(cuda_device).eval()
    xu, _ = synthetic_mixed_batch(6, 64, 96, seed=4)
    x_cpu = xu.permute(0, 3, 1, 2).float() / 255
    x = x_cpu.to(cuda_device)
    probs, _ = predict_probs(clf, x)
    p_cpu, _ = predict_probs(clf_cpu, x_cpu)
    assert float((probs.cpu() - p_cpu).abs().max()) <= 1e-3
    th = [float(probs[:, i].median()) for i in range(len(DEGRADATIONS))]
    pipe = MultiDegradationPipeline(clf, enh, thresholds=th)
    y = pipe(x)
    cols = [DEGRADATIONS.index(c) for c in ENHANCER_CLASSES]
    active = (pipe.last_probs[:, cols] >= torch.tensor([th[c] for c in cols], device=cuda_device)).cpu()
    assert 0 < int(active.sum()) < active.numel()
    with torch.no_grad():
        for i in range(x.shape[0]):
            t = x[i:i + 1].clone()
            for k, name in enumerate(ENHANCER_CLASSES):
                if active[i, k]:
                    t = enh[name](t)
            assert torch.equal(t[0], y[i]), i
