#!/bin/bash
# Quick GPU regression: op tests, tcgen05 tests, model tests, postproc, smoke, short bench.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tcgen05" --timeout 60 > gpurun_out/t_tc.txt 2>&1; echo "tcgen05 rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not tcgen05" --timeout 120 > gpurun_out/t_ops.txt 2>&1; echo "ops rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 > gpurun_out/t_model.txt 2>&1; echo "model rc=$?" >> gpurun_out/summary.txt
if [ "$1" != "nopost" ]; then timeout 600 python -m pytest tests/test_gpu_postproc.py -q -m gpu --timeout 300 > gpurun_out/t_post.txt 2>&1; echo "postproc rc=$?" >> gpurun_out/summary.txt; fi
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t_*.txt; do echo "== $f"; tail -n 3 $f; done
cat gpurun_out/bench.txt
tail -n 5 gpurun_out/bench.err
