"""Host logic of the multi-degradation router: bucketing, fixed class order, identity for unlabelled
images — with stub enhancers on CPU (the CDAN enhancers themselves are covered by the GPU test)."""
import pytest
import torch

from routing import DegradationRouter, active_classes


def test_thresholding_follows_reference_rule():
    p = torch.tensor([[0.5, 0.49], [0.2, 0.9]])
    assert active_classes(p, 0.5).tolist() == [[True, False], [False, True]]
    assert active_classes(p, [0.6, 0.4]).tolist() == [[False, True], [False, True]]
    with pytest.raises(ValueError):
        active_classes(p, [0.1, 0.2, 0.3])


def test_router_buckets_applies_in_order_and_keeps_unlabelled_images():
    calls = []

    def make(name, fn):
        def enh(x):
            calls.append((name, x.shape[0]))
            return fn(x)
        return enh

    router = DegradationRouter({"noise": make("noise", lambda x: x + 1.0), "blur": make("blur", lambda x: x * 2.0)},
                               class_order=["noise", "blur"], thresholds=0.5)
    x = torch.arange(4, dtype=torch.float32).view(4, 1, 1, 1).expand(4, 3, 2, 2).contiguous()
    probs = torch.tensor([[0.9, 0.1],    # noise only      -> x + 1
                          [0.1, 0.8],    # blur only       -> 2x
                          [0.7, 0.6],    # both, in order  -> 2(x + 1)
                          [0.0, 0.0]])   # none            -> identity
    y = router(x, probs)
    assert y[:, 0, 0, 0].tolist() == [1.0, 2.0, 6.0, 3.0]
    assert calls == [("noise", 2), ("blur", 2)] and router.last_bucket_sizes == {"noise": 2, "blur": 2}
    assert torch.equal(x[:, 0, 0, 0], torch.arange(4, dtype=torch.float32))  # the input batch is not modified


def test_router_argument_checks():
    with pytest.raises(KeyError):
        DegradationRouter({"noise": lambda x: x}, ["noise", "blur"])
    r = DegradationRouter({"noise": lambda x: x}, ["noise"])
    with pytest.raises(ValueError):
        r(torch.zeros(2, 3, 8, 8), torch.zeros(2, 2))
    bad = DegradationRouter({"noise": lambda x: x[:, :, :4]}, ["noise"])
    with pytest.raises(RuntimeError):
        bad(torch.zeros(1, 3, 8, 8), torch.ones(1, 1))
