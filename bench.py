#!/usr/bin/env python
"""bench.py — throughput of the detection hot path on B200 (contract: see DESIGN.md §Measurement).

Default workload = BASELINE.json configs[1]: YOLOv4 608x608, batch 64 per GPU, best-anchor target assignment
(GetTargets) + yolo_loss with the CIoU ignore mask (GetLoss), synthetic heads ~ N(0,1) and synthetic ground truth
(1..100 boxes/image).  One "step" = one pass of that path over one batch.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                      CPU restatement of the reference on the host cores

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F = np.float32
WORKLOADS = {
    # name: (image, per-GPU batch, description, algorithmic bytes / image per SURVEY.md §8d)
    "yolov4_608_b64_targets_loss": dict(image=608, batch=64, bytes_per_img=23197860,
                                        what="YOLOv4 608x608 batch 64: GetTargets + GetLoss(ciou ignore mask)"),
}
SEED = 20261018 + 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="yolov4_608_b64_targets_loss", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (0 = workload default)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2-fetch", type=int, default=0, help="set cudaLimitMaxL2FetchGranularity (32/64/128), 0 = leave")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step launch by launch instead of replaying a CUDA graph")
    ap.add_argument("--only-step", action="store_true", help="skip phase timing / e2e / cpu baseline (profiling runs)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the busiest samples are the ones under load
        sm_sorted = sorted(sm)
        top = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(wl, batch, rank):
    from tfmv_b200 import synth
    rng = np.random.default_rng(SEED + 1000 * rank)
    image = wl["image"]
    heads = synth.yolo_heads(rng, batch, image)
    boxes, classes, off = synth.gt_batch(rng, batch, (image, image), max_boxes=100)
    return heads, boxes, classes, off


# ------------------------------------------------------------------------------------------------
def cpu_step_images(args_tuple):
    """Oracle (NumPy restatement of the reference) on a few images: GetTargets + GetLoss.  Used by both the
    cpu_baseline leg and --impl reference; runs in worker processes."""
    image, seed, n_images = args_tuple
    from oracle import yolo as oy
    from tfmv_b200 import synth
    rng = np.random.default_rng(seed)
    anc = synth.yolo_anchors().astype(F)
    heads = synth.yolo_heads(rng, n_images, image)
    boxes, classes, off = synth.gt_batch(rng, n_images, (image, image), max_boxes=100)
    t0 = time.perf_counter()
    per = [oy.get_targets(boxes[off[b]:off[b + 1]], classes[off[b]:off[b + 1]], anc, (image, image), 80) for b in range(n_images)]
    y_true = [np.stack([p[l] for p in per], 0) for l in range(3)]
    loss = oy.get_loss(y_true, heads, (image, image), anc, 0.5, "ciou")
    return time.perf_counter() - t0, float(loss)


def run_reference(args, wl):
    """Reference arm: the oracle port on every host core (one image per worker per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_worker = 4
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        jobs = lambda s: [(wl["image"], SEED + 7919 * s + w, per_worker) for w in range(cores)]
        warm = max(1, min(args.warmup, 2))
        for w in range(warm):
            pool.map(cpu_step_images, jobs(10_000 + w))
        steps = max(1, min(args.steps, 30))
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(cpu_step_images, jobs(s))
        dt = time.perf_counter() - t0
    imgs = steps * cores * per_worker
    value = imgs / dt
    line = {
        "impl": "reference", "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["what"], "image": wl["image"], "per_gpu_batch": args.batch or wl["batch"],
                   "note": "NumPy restatement of the reference (TensorFlow not installable); each step = %d images, one per host core" % (cores * per_worker)},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d images (GetTargets+GetLoss ciou, 608x608), process pool over all host cores" % (steps, cores * per_worker)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import ctypes
    import torch
    import torch.distributed as dist
    import tfmv_b200  # noqa: F401
    from tfmv_b200 import _lib, _tensors as T, synth
    from tfmv_b200.ai_models.datasets.coco_dataset import DataGenerator
    from tfmv_b200.ai_models.utils import tf_yolo_utils as tyu

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback in the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if args.l2_fetch:
        _lib.check(lib.b200_set_l2_fetch_granularity(args.l2_fetch), "set_l2_fetch_granularity")
    batch = args.batch or wl["batch"]
    image = wl["image"]
    anc = synth.yolo_anchors().astype(F)
    heads_h, boxes_h, classes_h, off_h = make_inputs(wl, batch, rank)
    # pinned host copies (e2e path) and resident device copies (device-timed path)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    heads_p = [pin(h) for h in heads_h]
    boxes_p, classes_p, off_p = pin(boxes_h), pin(classes_h), pin(off_h)
    heads_d = [h.to(dev, non_blocking=True) for h in heads_p]
    boxes_d, classes_d, off_d = boxes_p.to(dev), classes_p.to(dev), off_p.to(dev)
    gen = DataGenerator(80, anc, (image, image))
    A, RF = 3, 85
    y_true = tuple(torch.empty((batch, hw[0], hw[1], A, RF), dtype=torch.float32, device=dev) for hw in gen.layers_hw)
    hw = (ctypes.c_int32 * 6)(*[d for l in gen.layers_hw for d in l])
    ws = torch.empty((lib.b200_yolo_loss_workspace_bytes(hw, batch, A),), dtype=torch.uint8, device=dev)
    global_batch = batch * world
    parts_buf = {}

    def log(msg):
        if args.verbose:
            sys.stderr.write("[rank %d] %s\n" % (rank, msg))
            sys.stderr.flush()

    def local_step(heads, boxes, classes, off):
        """This rank's images: target assignment + loss partials (already divided by the global batch)."""
        gen.GetTargetsBatch(classes, boxes, off, out=y_true)
        loss, parts = tyu._loss_call(y_true, heads, (image, image), anc, 0.5, "ciou", 0, batch_divisor=global_batch,
                                     return_parts=True, workspace=ws)
        return loss, parts

    def exchange(loss, parts):
        if world > 1:
            loss = tyu.combine_loss_parts(parts)  # the single collective of the path: 12 floats over NCCL
        parts_buf["loss"] = loss
        return loss

    def step(heads, boxes, classes, off):
        return exchange(*local_step(heads, boxes, classes, off))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    pending_streams = []   # side streams whose work belongs to the timed region

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        for st_ in pending_streams:
            torch.cuda.current_stream().wait_stream(st_)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- device-resident timing (the `value`) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    from tfmv_b200 import runtime
    log("inputs ready")
    if args.no_graph:
        dev_step = lambda: step(heads_d, boxes_d, classes_d, off_d)
    else:
        # the kernels of the step replay as one CUDA graph; the 12-float all-reduce is issued right behind it on the
        # same stream (kept outside the capture so the graph does not depend on NCCL's capture support)
        local_graph = runtime.capture(lambda: local_step(heads_d, boxes_d, classes_d, off_d))
        if world == 1:
            dev_step = lambda: exchange(*local_graph())
        else:
            # two graphs with their own output buffers: the all-reduce of step i runs on a communication stream
            # while the kernels of step i+1 already execute (it only needs the 12 floats step i produced)
            graphs = [local_graph, runtime.capture(lambda: local_step(heads_d, boxes_d, classes_d, off_d))]
            comm = torch.cuda.Stream()
            done = [None, None]
            counter = [0]

            def dev_step():
                k = counter[0] & 1
                counter[0] += 1
                main = torch.cuda.current_stream()
                if done[k] is not None:
                    main.wait_event(done[k])      # step i-2's exchange has released this buffer pair (long ago)
                loss, parts = graphs[k]()
                ready = torch.cuda.Event()
                ready.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ready)
                    exchange(loss, parts)
                    done[k] = torch.cuda.Event()
                    done[k].record(comm)

            pending_streams.append(comm)
    log("graph captured")
    if world > 1:
        step(heads_d, boxes_d, classes_d, off_d)  # NCCL communicator warm-up outside the timed region
        barrier()
        log("nccl warm")
    ms_dev = timed(dev_step, args.steps, max(args.warmup, 3))
    del pending_streams[:]
    log("device timing done")
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(parts_buf["loss"].item())

    if args.only_step:
        if rank == 0:
            print(json.dumps({"value": global_batch * args.steps / (ms_dev / 1e3), "ms_per_step": ms_dev / args.steps, "loss": loss_val}))
        return
    # ---- per-phase timing on the launching stream (roofline of the dominant kernel) ----
    st = T.stream_ptr()
    n_fill = sum(int(t.numel()) for t in y_true)
    tp = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in y_true])
    pp = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in heads_d])
    anc_h = np.ascontiguousarray(anc.reshape(-1))
    img_h = np.array([image, image], dtype=F)
    parts_t = torch.empty((3, 4), dtype=torch.float32, device=dev)
    loss_t = torch.empty((), dtype=torch.float32, device=dev)

    def ph_fill():  # the step's zero-fill: one fill_zero_multi_kernel launch (no boxes -> no scatter launch)
        lib.b200_yolo_assign_targets(boxes_d.data_ptr(), classes_d.data_ptr(), off_d.data_ptr(), batch, 0,
                                     anc_h.ctypes.data_as(ctypes.c_void_p), A, img_h.ctypes.data_as(ctypes.c_void_p), 80,
                                     hw, tp, 1, st)

    def ph_scatter():
        lib.b200_yolo_assign_targets(boxes_d.data_ptr(), classes_d.data_ptr(), off_d.data_ptr(), batch, boxes_d.shape[0],
                                     anc_h.ctypes.data_as(ctypes.c_void_p), A, img_h.ctypes.data_as(ctypes.c_void_p), 80,
                                     hw, tp, 0, st)

    def ph_loss_stage(mask):
        def run():
            lib.b200_yolo_loss_stages(tp, pp, hw, batch, A, 80, anc_h.ctypes.data_as(ctypes.c_void_p),
                                      img_h.ctypes.data_as(ctypes.c_void_p), 0.5, 2, 0, float(global_batch), parts_t.data_ptr(),
                                      loss_t.data_ptr(), ws.data_ptr(), ws.numel(), mask, st)
        return run

    n_rec = n_fill // RF
    ph_fill(); ph_scatter()
    phases = []
    # one entry per kernel of the step, each timed alone with CUDA events over `steps` back-to-back launches (the loss
    # kernels through the stage hook of the C ABI; their state lives in the workspace, so the order below matters)
    for name, fn, nbytes, launches in (
            ("fill_zero_multi_kernel (dense y_true zero-fill, write)", ph_fill, n_fill * 4, 1),
            ("yolo_scatter_targets_kernel (one CTA per image)", ph_scatter, int(boxes_d.shape[0]) * (16 + 4 + 340), 1),
            ("yolo_loss_scan_kernel (obj channel of y_true; dense-equivalent read of y_true)", ph_loss_stage(1), n_fill * 4, 1),
            ("yolo_loss_gtprep_kernel (a thread per object)", ph_loss_stage(2), int(boxes_d.shape[0]) * (16 + 32), 1),
            ("yolo_loss_ignore_kernel (box/conf logits of y_pred + object records; dense-equivalent read of y_pred)",
             ph_loss_stage(4), n_fill * 4, 1),
            ("yolo_loss_finalize_kernel (fp64 partial sums)", ph_loss_stage(8), n_rec // 128 * 8, 1)):
        ms = timed(fn, args.steps, 3) / args.steps
        phases.append({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbps": nbytes / ms / 1e6, "launches": launches})
        if fn in (ph_fill, ph_scatter):
            ph_fill(); ph_scatter()  # restore valid targets (repeated scatters collide with themselves)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    dom = max(phases, key=lambda p: p["ms"])
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = dom["kernel"].split(" ")[0].replace("_kernel", "")
        traffic = tj.get(key)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbps"], "peak": peak, "unit": "GB/s",
                "frac": dom["gbps"] / peak, "traffic": traffic, "peak_source": peak_src,
                "step_dense_equivalent_gbps": wl["bytes_per_img"] * batch / (ms_dev / args.steps) / 1e6,
                "phases": phases}
    if traffic:
        roofline["achieved_dram_gbps"] = traffic / dom["ms"] / 1e6
        roofline["frac_dram"] = traffic / dom["ms"] / 1e6 / peak
    roofline["peak_note"] = ("peak is the measured COPY bandwidth (read + write); the write-only zero-fill runs above it, and the "
                             "dense-equivalent figures of the sector-sparse loss kernels are not physical traffic")
    roofline["phases_note"] = ("each kernel timed alone, launched back to back from Python: entries below ~15 us are bounded by "
                               "the launch interval, their device durations are in profiles/r01_launches_step_v13.csv")
    roofline["loss_kernels_dense_equivalent_gbps"] = 2 * n_fill * 4 / sum(p["ms"] for p in phases[2:]) / 1e6
    if dom["kernel"].startswith("yolo_loss"):
        roofline["note"] = ("achieved/frac use SURVEY 8(d)'s dense algorithmic bytes (y_true + y_pred read once); the loss "
                            "kernels are sector-sparse (obj*(...) makes the class channels of non-object cells dead data), so "
                            "the dense-equivalent figure can exceed the physical peak; achieved_dram_gbps/frac_dram use the "
                            "ncu-measured DRAM bytes of the same launch (profiles/traffic.json)")
    log("phase timing done")
    # ---- same step with persistent target buffers (sparse reset instead of the dense zero-fill); reported beside
    # the headline, which keeps the reference's fresh-zeros-every-call behaviour ----
    persistent = None
    if world == 1:
        from tfmv_b200.ai_models.datasets.coco_dataset import TargetBuffers
        tbuf = TargetBuffers()

        def pstep():
            yt = gen.GetTargetsBatch(classes_d, boxes_d, off_d, buffers=tbuf)
            return tyu._loss_call(yt, heads_d, (image, image), anc, 0.5, "ciou", 0, batch_divisor=global_batch,
                                  return_parts=True, workspace=ws)
        pstep()  # first call: dense fill
        pfn = pstep if args.no_graph else runtime.capture(pstep)
        ms_p = timed(pfn, args.steps, 3)
        ploss = float(pfn()[0].item())
        persistent = {"what": "GetTargetsBatch(buffers=TargetBuffers) + GetLoss: y_true reused across steps, only the previous "
                              "step's records are re-zeroed", "value": global_batch * args.steps / (ms_p / 1e3),
                      "unit": "images/s", "ms_per_step": ms_p / args.steps, "loss": ploss, "loss_equal": ploss == loss_val}
    # ---- end-to-end through the public API with host buffers ----
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(lambda: float(step(heads_p, boxes_p, classes_p, off_p).item()), e2e_steps, 2)
    h2d = sum(h.numel() * 4 for h in heads_p) + boxes_p.numel() * 4 + classes_p.numel() * 4 + off_p.numel() * 4
    e2e = {"value": global_batch * e2e_steps / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(h2d) * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / e2e_steps}  # bytes: all ranks together

    # for comparison: the same call after an explicit copy of the whole y_pred to the device
    def copy_step():
        hd = [h.to(dev, non_blocking=True) for h in heads_p]
        return float(step(hd, boxes_p, classes_p, off_p).item())
    ms_cp = timed(copy_step, e2e_steps, 2)
    e2e["y_pred_transfer"] = ("pinned host y_pred is read in place by the loss kernels (they need ~6 % of it: 32-byte sectors "
                              "fetched over PCIe); boxes / classes / offsets are copied")
    n_rec_all = sum(int(h.numel()) for h in heads_p) // RF
    e2e["h2d_bytes_fetched_estimate"] = world * int(n_rec_all * 48 + boxes_p.shape[0] * 352 + boxes_p.numel() * 4 + classes_p.numel() * 4
                                                    + off_p.numel() * 4)  # ~1.5 32-byte sectors per record + the object records
    e2e["h2d_bytes_per_step_note"] = "size of the host tensors handed to the call; the kernels fetch only the sectors they use"
    e2e["explicit_copy"] = {"value": global_batch * e2e_steps / (ms_cp / 1e3), "unit": "images/s", "ms_per_step": ms_cp / e2e_steps,
                            "what": "whole y_pred copied host->device first (495 MB per step, PCIe-bound)"}

    # ---- CPU baseline beside it (rank 0, N == 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = args.cpu_sample or 32
        cpu_step_images((image, SEED + 5, 1))  # warm-up
        dt, _ = cpu_step_images((image, SEED + 6, n))
        cpu = {"value": n / dt, "unit": "images/s", "cores": 1, "kind": "port",
               "sample": "%d images of the same workload (NumPy oracle GetTargets+GetLoss, single process), %.1f s" % (n, dt)}

    # ---- software-pipelined drop-in step: the targets of batch i+1 are assigned on a second stream while the loss of
    # batch i runs (double-buffered y_true), as the reference's tf.data prefetch overlaps GetTargets with train_step ----
    pipelined = None
    if world == 1 and not args.no_graph:
        y_true2 = tuple(torch.empty_like(t) for t in y_true)
        bufs = (y_true, y_true2)
        side = torch.cuda.Stream()               # target assignment (bandwidth-bound fill: takes whatever is left)
        hi = torch.cuda.Stream(priority=-1)      # loss kernels (latency-bound): their CTAs are scheduled first

        def make_pipe(k):
            def run():
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                hi.wait_stream(main)
                with torch.cuda.stream(side):
                    gen.GetTargetsBatch(classes_d, boxes_d, off_d, out=bufs[k ^ 1])   # batch i+1
                with torch.cuda.stream(hi):
                    out = tyu._loss_call(bufs[k], heads_d, (image, image), anc, 0.5, "ciou", 0, batch_divisor=global_batch,
                                         return_parts=True, workspace=ws)           # batch i
                main.wait_stream(side)
                main.wait_stream(hi)
                return out
            return run
        gen.GetTargetsBatch(classes_d, boxes_d, off_d, out=bufs[0])
        gen.GetTargetsBatch(classes_d, boxes_d, off_d, out=bufs[1])
        pipes = [runtime.capture(make_pipe(0)), runtime.capture(make_pipe(1))]
        pk = [0]

        def pipe_step():
            r = pipes[pk[0] & 1]()
            pk[0] += 1
            return r
        ms_pl = timed(pipe_step, args.steps, 4)
        plloss = float(pipe_step()[0].item())
        pipelined = {"what": "GetTargets(batch i+1) on a second stream under GetLoss(batch i), double-buffered dense y_true",
                     "value": global_batch * args.steps / (ms_pl / 1e3), "unit": "images/s", "ms_per_step": ms_pl / args.steps,
                     "loss": plloss, "loss_equal": plloss == loss_val}

    # ---- sparse-target fusion: the same step without materialising y_true ----
    fused = None
    if world == 1:
        ws_f = torch.empty((lib.b200_yolo_loss_from_boxes_workspace_bytes(hw, batch, A, int(boxes_d.shape[0])),), dtype=torch.uint8, device=dev)

        def fstep():
            return tyu.GetLossFromBoxes(classes_d, boxes_d, off_d, heads_d, (image, image), anc, 80, 0.5, "ciou",
                                        batch_divisor=global_batch, return_parts=True, workspace=ws_f)
        ffn = fstep if args.no_graph else runtime.capture(fstep)
        ms_f = timed(ffn, args.steps, 3)
        floss = float(ffn()[0].item())
        fused = {"what": "GetLossFromBoxes: target assignment + loss from the box lists, no dense y_true (API extension, SURVEY 8f N3)",
                 "value": global_batch * args.steps / (ms_f / 1e3), "unit": "images/s", "ms_per_step": ms_f / args.steps,
                 "loss": floss, "loss_rel_diff": abs(floss - loss_val) / abs(loss_val)}

    if rank == 0:
        line = {
            "metric": "images/sec", "value": global_batch * args.steps / (ms_dev / 1e3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["what"], "image": image, "per_gpu_batch": batch, "global_batch": global_batch,
                       "classes": 80, "anchors_per_cell": 3, "gt_boxes_per_image": "U{1..100}",
                       "parallelism": "dp%d (images sharded, one 12-float NCCL all-reduce per step%s)" % (
                           world, ", overlapped with the next step's kernels on a second stream" if world > 1 and not args.no_graph else ""),
                       "launch": "launch by launch" if args.no_graph else "CUDA graph replay of the step",
                       "l2": "inputs larger than L2 (y_pred %.0f MB + y_true %.0f MB per step vs 126 MB L2)" % (
                           n_fill * 4 / 1e6, n_fill * 4 / 1e6)},
            "loss": loss_val, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "persistent_targets": persistent, "sparse_target_fusion": fused, "pipelined_streams": pipelined,
            "gpu_launches": 6 * args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
