"""GPU parity: the serving view's box post-processing (views/object_detection.py:70-85) vs the oracle."""
import numpy as np
import pytest

from test_gpu_core import _t

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.mark.parametrize("old_wh,padding", [((640, 480), (52, 52, 0, 0)), ((333, 777), (0, 0, 119, 119)), ((416, 416), (0, 0, 0, 0)),
                                            ((1920, 1080), (91, 91, 0, 0))])
def test_restore_predictions_matches_oracle(lib, cuda, old_wh, padding):
    from oracle import views as ov
    from tfmv_b200.views.object_detection import restore_predictions
    rng = np.random.default_rng(20261018 + old_wh[0])
    n = 700
    c = rng.uniform(-0.1, 1.1, (n, 2))
    wh = np.exp(rng.uniform(np.log(1e-3), np.log(0.8), (n, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(F)
    boxes[::50] = boxes[1::50]                       # duplicates
    boxes[5, 2] = boxes[5, 0] + F(2.0 / old_wh[0])   # a width right at the 2 px limit
    cid = rng.integers(0, 80, n).astype(np.int32)
    sc = rng.random(n).astype(F)
    cl = rng.random((n, 80)).astype(F)
    cf = rng.random((n, 1)).astype(F)
    want = ov.restore_predictions(boxes, cid, sc, cl, cf, (416, 416), padding, old_wh)
    got = restore_predictions(_t(boxes, cuda), _t(cid, cuda), _t(sc, cuda), _t(cl, cuda), _t(cf, cuda), np.int32([416, 416]), padding,
                              np.int32(old_wh))
    assert 0 < want[0].shape[0] < n
    assert str(got[0].dtype) == "torch.int32"
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    # an empty input stays empty
    e = restore_predictions(np.zeros((0, 4), F), np.zeros((0,), np.int32), np.zeros((0,), F), np.zeros((0, 80), F), np.zeros((0, 1), F),
                            (416, 416), padding, old_wh)
    assert e[0].shape[0] == 0 and e[3].shape[0] == 0
