#!/bin/bash
# One-shot GPU validation: per-layer diagnostic, op tests, tcgen05 GEMM tests, model tests, bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gpu_diag.py rny002_gsf > gpurun_out/diag_rny002.txt 2>&1; echo "diag rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not tcgen05" --timeout 120 > gpurun_out/t_ops.txt 2>&1; echo "ops rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_postproc.py -q -m gpu --timeout 300 > gpurun_out/t_post.txt 2>&1; echo "postproc rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tcgen05" --timeout 60 > gpurun_out/t_tc.txt 2>&1; echo "tcgen05 rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 > gpurun_out/t_model.txt 2>&1; echo "model rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_ops.txt gpurun_out/t_post.txt gpurun_out/t_tc.txt gpurun_out/t_model.txt gpurun_out/smoke.txt
cat gpurun_out/bench.txt
