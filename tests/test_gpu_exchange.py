"""The loss all-reduce of the data-parallel path: protocol self-test of the
NVLink peer-mailbox exchange on one device (the ranks are CTAs of one grid, so nothing waits on a kernel that might not
be running), and — when the box has two or more GPUs — the real thing: one process per GPU, CUDA-IPC mailboxes,
GetLossSharded with the exchange fused into the finalize kernel and with the library's NCCL transport, against the
oracle on the whole batch."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


@pytest.mark.parametrize("world,n", [(2, 12), (4, 11), (8, 32), (3, 1)])
def test_peer_exchange_protocol_selftest(lib, cuda, world, n):
    import torch
    from tfmv_b200 import _lib, _tensors as T
    rounds = 64
    ws = torch.empty((world * lib.b200_peer_mailbox_bytes(),), dtype=torch.uint8, device=cuda)
    out = torch.full((world, rounds, n), -1.0, dtype=torch.float32, device=cuda)
    _lib.check(lib.b200_peer_exchange_selftest(world, rounds, n, ws.data_ptr(), ws.numel(), out.data_ptr(), T.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    q = np.arange(1, world + 1, dtype=np.float64)[:, None, None]
    i = np.arange(1, n + 1, dtype=np.float64)[None, None, :]
    k = np.arange(rounds, dtype=np.float64)[None, :, None]
    want = (q * i + k).sum(0)                       # every rank sees the same sums
    got = out.cpu().numpy()
    for r in range(world):
        assert np.array_equal(got[r], want.astype(F)), r
    # header of every mailbox: `rounds` exchanges completed, no timeouts
    hdr = ws.view(world, -1)[:, :16].cpu().numpy().view(np.uint64)
    assert (hdr[:, 0] == rounds).all() and (hdr[:, 1] & 0xffffffff == 0).all()


def test_exchange_argument_checks(lib, cuda):
    import torch
    t = torch.zeros(12, device=cuda)
    boxes = (ctypes.c_void_p * 8)()
    assert lib.b200_allreduce_loss_peer(t.data_ptr(), 12, 0, 1, None, None) == 0       # world 1: nothing to do
    assert lib.b200_allreduce_loss_peer(t.data_ptr(), 12, 0, 9, boxes, None) == -1     # more than one NVSwitch domain
    assert lib.b200_allreduce_loss_peer(t.data_ptr(), 12, 2, 2, boxes, None) == -1     # rank outside the world
    assert lib.b200_allreduce_loss_peer(t.data_ptr(), 12, 0, 2, boxes, None) == -1     # unmapped mailbox
    assert b"mailbox of rank" in lib.b200_last_error()
    assert lib.b200_allreduce_loss_peer(t.data_ptr(), 33, 0, 2, boxes, None) == -1
    assert lib.b200_allreduce_loss(None, t.data_ptr(), 12, None) != 0                  # NULL communicator


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        import tfmv_b200  # noqa: F401
        from oracle import yolo as oy
        from tfmv_b200 import runtime, synth
        from tfmv_b200.ai_models.utils.tf_yolo_utils import GetLossSharded, shard_range
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        rng = np.random.default_rng(77)
        image, batch = 128, 6
        anc = synth.yolo_anchors().astype(F)
        boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=12)
        per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc / F(image), (image, image), 80) for b in range(batch)]
        y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
        y_pred = synth.yolo_heads(rng, batch, image)
        lo, hi = shard_range(batch, rank, world)
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        yt, yp = [to(t[lo:hi]) for t in y_true], [to(t[lo:hi]) for t in y_pred]
        want = float(oy.get_loss(y_true, y_pred, (image, image), anc, 0.5, "ciou"))
        res = {"want": want}
        peer = runtime.PeerExchange()
        # eager, then the same call replayed from a CUDA graph (the exchange is part of the captured step)
        res["peer"] = float(GetLossSharded(yt, yp, (image, image), anc, 0.5, "ciou", global_batch=batch, exchange=peer))
        step = runtime.capture(lambda: GetLossSharded(yt, yp, (image, image), anc, 0.5, "ciou", global_batch=batch, exchange=peer))
        vals = [float(step()) for _ in range(5)]
        res["peer_graph"] = vals
        res["peer_status"] = peer.status()
        # split exchange: publish inside the loss call, collect on a second stream one step later (bench.py's N > 1 mode)
        from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu
        comm = torch.cuda.Stream()
        outs = [dict(parts=torch.empty((3, 4), device=dev), loss=torch.zeros((), device=dev), done=None) for _ in range(2)]
        got = []
        for i in range(6):
            k = i & 1
            main = torch.cuda.current_stream()
            if outs[k]["done"] is not None:
                main.wait_event(outs[k]["done"])          # publish(f) behind the own collect(f - 2)
                got.append(float(outs[k]["loss"]))
            tyu._loss_call(yt, yp, (image, image), anc, 0.5, "ciou", 0, batch_divisor=batch, exchange=peer, defer_collect=True)
            ready = torch.cuda.Event(); ready.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                peer.collect_yolo(outs[k]["parts"], outs[k]["loss"])
                outs[k]["done"] = torch.cuda.Event(); outs[k]["done"].record(comm)
        torch.cuda.synchronize()
        got += [float(outs[0]["loss"]), float(outs[1]["loss"])]
        res["peer_deferred"] = got
        res["peer_status"] = peer.status()
        nccl = runtime.NcclExchange()
        res["nccl"] = float(GetLossSharded(yt, yp, (image, image), anc, 0.5, "ciou", global_batch=batch, exchange=nccl))
        res["torch"] = float(GetLossSharded(yt, yp, (image, image), anc, 0.5, "ciou", global_batch=batch))
        dist.barrier()
        torch.cuda.synchronize()
        nccl.close()
        peer.close()
        dist.destroy_process_group()
        q.put((rank, res))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, {"error": "%s\n%s" % (e, traceback.format_exc())}))


def test_sharded_loss_over_peer_mailboxes_and_nccl(lib, cuda):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs two or more GPUs (one process per GPU)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert "error" not in res[r], res[r]["error"]
    want = res[0]["want"]
    for r in range(world):
        v = res[r]
        assert abs(v["peer"] - want) <= 1e-4 * abs(want), v
        assert v["peer"] == res[0]["peer"]                       # rank-ordered sum: identical bits on every rank
        assert all(x == v["peer"] for x in v["peer_graph"]), v   # graph replays reproduce the eager value
        assert v["peer_status"][1] == 0 and v["peer_status"][0] >= 12
        assert all(x == v["peer"] for x in v["peer_deferred"]), v   # the split exchange gives the same bits
        assert abs(v["nccl"] - want) <= 1e-4 * abs(want) and abs(v["torch"] - want) <= 1e-4 * abs(want)
