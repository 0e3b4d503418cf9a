#!/bin/bash
# merged exact pass: parity + step time against the single-kernel form
mkdir -p gpurun_out
python -m pytest tests/test_gpu_yolo_loss.py tests/test_gpu_exchange.py tests/test_gpu_fullsize_properties.py tests/test_golden.py -m gpu -x -q > gpurun_out/finx_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/finx_pytest.log
for v in "10:592" "10:296" "10:1184" "12:592" "8:592" "0:0"; do
  s=${v%%:*}; c=${v##*:}
  echo "== split=$s ctas=$c" >> gpurun_out/finx_bench.log
  B200_YL_SPLIT=$s B200_YL_FINX_CTAS=$c python bench.py --only c2 --only-step --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" >> gpurun_out/finx_bench.log 2>&1
done
for s in 10 0; do
B200_YL_SPLIT=$s python bench.py --only c5 --c5-global-batch 64 --only-step --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); c=d['configs']['c5']; print('c5 b64 split=$s', c.get('ms_per_step'), c.get('value'))" >> gpurun_out/finx_bench.log 2>&1
done
tail -3 gpurun_out/finx_pytest.log; cat gpurun_out/finx_bench.log
