"""SURVEY §8(b) threading contract: the reference calls these functions from parallel tf.data map workers
(datasets/coco_dataset.py:328) and TF executor threads, so the C ABI must be re-entrant — no global mutable state, an
explicit stream per call, a thread-local error slot.  Four host threads, each on its own CUDA stream, drive GetTargets,
GetLoss and GetNMSBoxes concurrently; every result must be bit-identical to the same calls issued serially."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


def _job(k, cuda):
    """Inputs of worker k (different sizes per worker so that workspaces and grids differ)."""
    import torch
    from tfmv_b200 import synth
    rng = np.random.default_rng(900 + k)
    image = (128, 160, 192, 224)[k % 4]
    batch = 2 + k % 3
    heads = [torch.from_numpy(h).to(cuda) for h in synth.yolo_heads(rng, batch, image)]
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=25)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    return dict(image=image, batch=batch, heads=heads, boxes=to(boxes), classes=to(classes), off=to(off))


def _run(job, rounds):
    """GetTargets -> GetLoss(ciou) -> GetNMSBoxes(diou) `rounds` times on the CURRENT stream; returns host copies."""
    import torch
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu
    anc = synth.yolo_anchors().astype(F)
    image = job["image"]
    gen = DataGenerator(80, anc / F(image), (image, image))
    out = None
    for _ in range(rounds):
        y_true = gen.GetTargetsBatch(job["classes"], job["boxes"], job["off"])
        loss = tyu.GetLoss(y_true, job["heads"], (image, image), anc, 0.5, "ciou")
        r = tyu.GetNMSBoxesBatch(*job["heads"], anc, (image, image), 80, 0.5, 0.3, 0.5, "diou", with_indices=True)
        out = (loss, y_true, r)
    torch.cuda.current_stream().synchronize()
    loss, y_true, r = out
    cnt = r["count"].cpu().numpy()
    return dict(loss=float(loss), y_true=[t.cpu().numpy() for t in y_true], count=cnt,
                sel=[r["sel_idx"][b, :cnt[b]].cpu().numpy() for b in range(len(cnt))],
                boxes=[r["boxes"][b, :cnt[b]].cpu().numpy() for b in range(len(cnt))],
                classes=[r["classes"][b, :cnt[b]].cpu().numpy() for b in range(len(cnt))])


def test_four_threads_four_streams_equal_serial(lib, cuda):
    import torch
    n = 4
    jobs = [_job(k, cuda) for k in range(n)]
    serial = [_run(j, 1) for j in jobs]
    results, errors = [None] * n, []
    start = threading.Barrier(n)

    def worker(k):
        try:
            torch.cuda.set_device(cuda)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.default_stream())
            with torch.cuda.stream(s):
                start.wait()
                results[k] = _run(jobs[k], 25)   # many rounds: the calls of the four threads interleave on the device
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    for k in range(n):
        a, b = serial[k], results[k]
        assert b is not None
        assert a["loss"] == b["loss"], k                                   # bit-reproducible loss (sorted object lists)
        assert all(np.array_equal(x, y) for x, y in zip(a["y_true"], b["y_true"])), k
        assert np.array_equal(a["count"], b["count"]), k
        for key in ("sel", "boxes", "classes"):
            assert all(np.array_equal(x, y) for x, y in zip(a[key], b[key])), (k, key)


def test_error_slot_is_thread_local(lib):
    """b200_last_error(): a failing call in one thread must not overwrite the message another thread reads."""
    import ctypes
    seen = {}

    def bad(name, fn):
        fn()
        seen[name] = lib.b200_last_error()

    t1 = threading.Thread(target=bad, args=("a", lambda: lib.b200_set_l2_fetch_granularity(48)))
    t1.start(); t1.join()
    assert lib.b200_pairwise_iou(0, 5, 0, 5, 0, 0, 0) == -1
    mine = lib.b200_last_error()
    t2 = threading.Thread(target=bad, args=("b", lambda: lib.b200_set_l2_fetch_granularity(7)))
    t2.start(); t2.join()
    assert lib.b200_last_error() == mine and b"pairwise" in mine
    assert b"48" in seen["a"] and b"7" in seen["b"]
