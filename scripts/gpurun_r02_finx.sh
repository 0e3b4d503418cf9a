#!/bin/bash
# scan tail (gtprep folded in) + reverse scan: parity + step time
mkdir -p gpurun_out; rm -f gpurun_out/finx_bench.log
python -m pytest tests/test_gpu_yolo_loss.py tests/test_gpu_exchange.py tests/test_gpu_fullsize_properties.py tests/test_golden.py tests/test_gpu_reentrancy.py -m gpu -x -q > gpurun_out/finx_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/finx_pytest.log
run() { echo "== $*" >> gpurun_out/finx_bench.log; env "$@" python bench.py --only c2 --only-step --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" >> gpurun_out/finx_bench.log 2>&1; }
run A=1
run B200_YL_SCAN_REV=0
run B200_YL_GTPREP_LAUNCH=1
run B200_YL_SCAN_REV=0 B200_YL_GTPREP_LAUNCH=1
run A=2
tail -3 gpurun_out/finx_pytest.log; cat gpurun_out/finx_bench.log
