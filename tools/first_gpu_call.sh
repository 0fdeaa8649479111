#!/bin/bash
# One GPU call that (re)validates what round 1 could not run any more, meant as the first `gpurun` of the next round:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# 1. the whole GPU suite (the late additions of tests/test_zz_late_additions_gpu.py ran only step by step so far)
# 2. the never-run TMA bulk-copy variant of the tiled SpMM, under its own timeout (the kernel traps instead of hanging)
# 3. its effect on the C4 SpMM numbers, and the contour step at 16 / 128 nodes as the latency / throughput baseline
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -8) | tee gpurun_out/r2_first_pytest.log
(NEPB_RUN_UNVALIDATED=1 timeout 120 python -m pytest tests/test_zz_late_additions_gpu.py -q -k tma 2>&1 | tail -5) | tee gpurun_out/r2_first_tma.log
(timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | grep "\[bench\]") | tee gpurun_out/r2_first_bench_default.log
(NEPB_SPMM_BULK=1 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | grep "spmm k=8") | tee gpurun_out/r2_first_bench_bulk.log
timeout 60 python tools/contour_step.py 3 10 16 16 | tee gpurun_out/r2_first_step16.log
timeout 60 python tools/contour_step.py 3 5 128 128 | tee gpurun_out/r2_first_step128.log
(timeout 120 python tools/qdep0_lu_diag.py 2>&1 | tail -40) | tee gpurun_out/r2_first_qdep0.log
