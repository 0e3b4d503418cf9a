"""CPU: oracle/views.py (NumPy-1.x casting emulated) against lines 71-85 of the reference's serving view executed
verbatim (tests/golden/make_golden_views.py -> tests/golden/ref_views.npz)."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_views import CASES, inputs  # noqa: E402

from oracle import views as ov  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_views.npz"))
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_reproduces_the_reference_view_lines(i):
    old_wh, padding = CASES[i]
    b, cid, sc, cl, cf = inputs(i, old_wh)
    boxes, g_cid, g_sc, g_cl, g_cf = ov.restore_predictions(b, cid, sc, cl, cf, (416, 416), padding, old_wh)
    want = GOLD["%d/y_boxes" % i]
    assert 0 < want.shape[0] < b.shape[0] and boxes.dtype == np.int32
    assert np.array_equal(boxes, want)
    assert np.array_equal(g_cid, GOLD["%d/y_classes_id" % i]) and np.array_equal(g_sc, GOLD["%d/y_scores" % i])
    assert sha(g_cl) == str(GOLD["%d/y_classes_sha" % i]) and sha(g_cf) == str(GOLD["%d/y_confidence_sha" % i])
