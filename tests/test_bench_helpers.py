"""CPU tests of the measurement plumbing: bench.py's helpers and the ncu launch-list summariser that feeds
`roofline.traffic` / `frac_dram` / the dominant kernel (profiles/r02_kernels.json)."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = [path]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_bench_helpers():
    b = _load(os.path.join(ROOT, "bench.py"), "bench_mod")
    assert b.auto_graph_steps(20, 0) == 20 and b.auto_graph_steps(50, 0) == 25 and b.auto_graph_steps(7, 0) == 7
    assert b.auto_graph_steps(20, 4) == 4 and b.auto_graph_steps(20, 3) == 1       # a request that does not divide the window: per-step graphs
    s = b.spread([3.0, 1.0, 2.0, 10.0, 4.0])
    assert s["n"] == 5 and s["median"] == 3.0 and s["min"] == 1.0 and s["max"] == 10.0
    assert sorted(b.BYTES) == sorted(b.WHAT) == ["c1", "c2", "c3", "c4", "c5"]
    assert b.BYTES["c2"] == 3 * 7732620 and b.BYTES["c4"] == 441936 * (81 * 4 + 32)


def test_launch_list_summariser_reproduces_the_committed_table():
    s = _load(os.path.join(ROOT, "profiles", "summarize_launches.py"), "summarize_mod")
    table = json.load(open(os.path.join(ROOT, "profiles", "r02_kernels.json")))
    for cfg in ("c2", "c3", "c4"):
        t = table[cfg]
        got = s.summarize(os.path.join(ROOT, "profiles", t["source"]), t["calls"], t["batch"])
        assert got["dominant"]["name"] == t["dominant"]["name"]
        assert abs(got["dram_bytes_per_step"] - t["dram_bytes_per_step"]) < 1.0
        assert abs(got["us_per_step_sum"] - t["us_per_step_sum"]) < 1e-6
        assert all(k["launches_per_step"] >= 0.9 for k in got["kernels"])                 # input preparation is listed apart
        assert abs(sum(k["share"] for k in got["kernels"]) - 1.0) < 1e-9
    # the kernels of the headline step are exactly the five of DESIGN.md section 4 (GT preparation rides in the scan
    # kernel, the exact ignore pass in the finalize launch)
    names = sorted(k["name"].split("(")[0].replace("void ", "").split("<")[0] for k in table["c2"]["kernels"])
    assert names == ["fill_zero_multi_kernel", "yolo_loss_finalize_kernel", "yolo_loss_ignore_lean_kernel",
                     "yolo_loss_scan_kernel", "yolo_scatter_targets_kernel"]
