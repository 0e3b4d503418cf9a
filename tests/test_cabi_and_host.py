"""CPU tests: the C-ABI library loads and exports every symbol include/b200det.h declares; host-side logic of the
shims (argument checks, error behaviour, no CPU fallback); data-parallel exchange step on gloo (world_size 2)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200det.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from tfmv_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libb200det.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding missing for %s" % n
    assert sorted(_lib.SIGNATURES) == names  # and nothing is bound that the header does not declare
    assert lib.b200_version() == 200


def test_no_cpu_fallback_without_device(lib):
    import torch
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOU
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.b200_device_ok() != 0 and b"no CPU fallback" in lib.b200_last_error()
    with pytest.raises(RuntimeError):
        GetIOU(np.zeros((1, 1, 4), F), np.zeros((1, 1, 4), F))


def test_bad_arguments_are_reported_not_crashed(lib):
    import ctypes
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    assert lib.b200_nms(0, 0, 0, 0, 0, 1, 99, 0, ctypes.c_float(0.5), 0, ctypes.c_float(0), 10, 0, 0, 0) == -1
    assert b"bad metric" in lib.b200_last_error()
    assert lib.b200_nms(0, 0, 0, 0, 0, 1, 0, 0, ctypes.c_float(0.5), 0, ctypes.c_float(0), 100000, 0, 0, 0) == -4
    assert lib.b200_pairwise_iou(0, 5, 0, 5, 0, 0, 0) == -1
    assert lib.b200_set_l2_fetch_granularity(48) == -1
    hw = (ctypes.c_int32 * 6)(13, 13, 26, 26, 52, 52)
    assert lib.b200_yolo_decode_nms_workspace_bytes(hw, 2, 3, 500) > 2 * 10647 * 32
    assert lib.b200_yolo_loss_workspace_bytes(hw, 2, 3) > 0
    hw5 = (ctypes.c_int32 * 10)(64, 64, 32, 32, 16, 16, 8, 8, 4, 4)
    assert lib.b200_effdet_table_floats(5, hw5, 9) == 2 * (64 + 32 + 16 + 8 + 4) + 5 * 18
    # the entry points added for the extensions validate the same way
    f = ctypes.c_float
    assert lib.b200_yolo_loss_from_boxes_workspace_bytes(hw, 2, 3, 100) > lib.b200_yolo_loss_workspace_bytes(hw, 2, 3)
    assert lib.b200_yolo_loss_from_boxes(0, 0, 0, 0, 0, 0, hw, 2, 3, 80, 0, 0, f(0.5), 2, 0, f(2), 0, 0, 0, 0, 0, 0) == -1
    assert lib.b200_yolo_loss_stages(0, 0, hw, 2, 3, 80, 0, 0, f(0.5), 2, 0, f(2), 0, 0, 0, 0, 0, 0) == -1
    assert b"stages" in lib.b200_last_error()
    assert lib.b200_yolo_reset_targets(0, 0, 2, 10, 0, 3, 0, 80, hw, 0, 0) == -1
    assert lib.b200_unletterbox_boxes(0, 0, 2, 10, 0, 0, 0, 0, 0, 0, 0) == -1
    assert lib.b200_letterbox_image(0, 480, 640, 3, 416, 416, 0, 0, 0, 0, 0, 0) == -1
    assert lib.b200_effdet_assign_targets_indexed(5, hw5, 9, 0, 81, 2, 0, 0, 0, f(0.5), 0, 0, 0, 0) == -1
    assert lib.b200_focal_box_partial_sums_indexed(5, 0, 81, 0, 0, 0, 0, 0, f(0.25), f(1.5), f(0.1), f(0), 0, 0, 0, 0) == -1


def test_reference_assertions_on_iou_type():
    from tfmv_b200.ai_models.efficientnet.utils.iou import get_iou
    from tfmv_b200.ai_models.utils.tf_iou_utils import GetIOU, GetIOUNMS
    from tfmv_b200.ai_models.utils.tf_yolo_utils import GetLoss
    z = np.zeros((1, 1, 4), F)
    with pytest.raises(AssertionError):   # tf_iou_utils.py:18
        GetIOU(z, z, "giou")
    with pytest.raises(AssertionError):   # efficientnet/utils/iou.py:76
        get_iou(z, z, "xiou")
    with pytest.raises(AssertionError):
        GetIOUNMS(z[0], np.zeros(1, F), 10, iou_type="giou")
    with pytest.raises(AssertionError):
        GetLoss([z] * 3, [z] * 3, (416, 416), np.zeros((3, 3, 2)), iou_type="giou")


def test_anchor_table_and_feat_sizes_host_logic():
    from oracle import effdet as oe
    from tfmv_b200.ai_models.efficientnet.utils.anchors import Anchors
    from tfmv_b200.ai_models.efficientnet.utils.get_feat_sizes import get_feat_sizes
    assert get_feat_sizes((512, 512), 7) == oe.get_feat_sizes((512, 512), 7)
    assert get_feat_sizes((600, 300), 5) == oe.get_feat_sizes((600, 300), 5)
    args = (3, 7, (512, 512), 3, [(1.0, 1.0), (1.4, 0.7), (0.7, 1.4)], 4.0)
    a, o = Anchors(*args), oe.Anchors(*args)
    t = a._table_host
    off = 0
    for l, (h, w) in enumerate(a._level_hw):
        yc, xc = t[off:off + h], t[off + h:off + h + w]
        hy, hx = t[off + h + w:off + h + w + 9], t[off + h + w + 9:off + h + w + 18]
        off += h + w + 18
        want = o.boxes[l]
        got = np.stack([yc[:, None, None] - hy[None, None, :] + 0 * xc[None, :, None],
                        xc[None, :, None] - hx[None, None, :] + 0 * yc[:, None, None],
                        yc[:, None, None] + hy[None, None, :] + 0 * xc[None, :, None],
                        xc[None, :, None] + hx[None, None, :] + 0 * yc[:, None, None]], -1).astype(F)
        assert np.array_equal(got, want)
    with pytest.raises(AssertionError):
        Anchors(3, 7, (512, 512), 3, [(1.0, 1.0)], [4.0, 4.0])   # one anchor_scale per level (anc:36)


def test_shard_range_partitions_the_batch():
    from tfmv_b200.ai_models.utils.tf_yolo_utils import shard_range
    for batch, world in [(512, 8), (10, 4), (3, 8), (64, 1)]:
        r = [shard_range(batch, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == batch and all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    import tfmv_b200  # noqa: F401
    from oracle import yolo as oy
    from tfmv_b200 import synth
    from tfmv_b200.ai_models.utils.tf_yolo_utils import combine_loss_parts, shard_range
    rng = np.random.default_rng(11)
    image, batch = 96, 6
    anc = (synth.yolo_anchors().astype(F) / F(96)).astype(F)
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=8)
    per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc, (image, image), 80) for b in range(batch)]
    y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
    y_pred = synth.yolo_heads(rng, batch, image)
    lo, hi = shard_range(batch, rank, world)
    # what each rank's kernel produces: its shard's sums divided by the GLOBAL batch
    _, parts = oy.get_loss([t[lo:hi] for t in y_true], [t[lo:hi] for t in y_pred], (image, image), anc, 0.5, "ciou", return_parts=True)
    parts = torch.from_numpy(parts * F(hi - lo) / F(batch))
    loss = combine_loss_parts(parts)
    want = oy.get_loss(y_true, y_pred, (image, image), anc, 0.5, "ciou")
    q.put((rank, float(loss), float(want)))
    dist.destroy_process_group()


def test_sharded_loss_exchange_on_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len({round(r[1], 3) for r in res}) == 1          # both ranks hold the same loss after the all-reduce
    for _, got, want in res:
        assert abs(got - want) <= 1e-4 * abs(want)          # 1e-4 also after the all-reduce


def test_product_package_never_touches_the_oracle_or_the_reference():
    """The oracle is the checker, not the product: nothing under the package (Python or CUDA sources) may import,
    execute or read oracle/ or /root/reference; importing every product module must not pull `oracle` in either."""
    import re
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "tensorflow2-machine-vision_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("build", "__pycache__"):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(d, f), encoding="utf-8").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "/root/reference" in text.replace("`/root/reference`", ""):
                bad.append(os.path.relpath(os.path.join(d, f), ROOT))
    assert not bad, bad
    code = ("import sys, pkgutil, importlib; sys.path.insert(0, %r); import tfmv_b200\n"
            "for m in pkgutil.walk_packages(tfmv_b200.__path__, 'tfmv_b200.'):\n"
            "    if not m.name.endswith(('.build', '.libb200det')): importlib.import_module(m.name)\n"
            "assert not [k for k in sys.modules if k == 'oracle' or k.startswith('oracle.')], 'oracle imported by the product'\n"
            "print('ok')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
