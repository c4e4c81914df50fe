#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): tests, smoke, bench, ncu launch list, DRAM traffic
# of the GEMM launches of one step, and a full capture of the dominant kernel.  Outputs land in
# gpurun_out/ ; the summaries are copied to profiles/ afterwards (tools/summarise_launches.py,
# tools/gemm_traffic.py).
set -u
O=gpurun_out
export EDB_SPIN_TIMEOUT_MS=30000
timeout 600 python -m pytest tests -q -m gpu > $O/r02_final_pytest_gpu.log 2>&1; echo "pytest_exit=$?" >> $O/r02_final_pytest_gpu.log
tail -3 $O/r02_final_pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/r02_final_smoke.log 2>&1; tail -2 $O/r02_final_smoke.log
timeout 400 python bench.py > $O/r02_final_bench_1gpu.json 2> $O/r02_final_bench_1gpu.err; echo "bench_exit=$?"
cut -c1-600 $O/r02_final_bench_1gpu.json
timeout 120 python bench.py --impl torch-nccl --steps 10 --warmup 3 2>/dev/null | tail -1 > $O/r02_final_bench_1gpu_torch_nccl.json; cat $O/r02_final_bench_1gpu_torch_nccl.json
# launch list of one eager step (the CUDA graph replays the same launches)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file $O/r02_final_launches.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-parity \
    > $O/r02_bench_under_ncu.log 2>&1; echo "ncu_list=$?"
# DRAM bytes of every GEMM launch of the last step (-> roofline.traffic)
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_gemm_bf16 -s 1164 -c 291 --csv --log-file $O/r02_final_gemm_dram.csv \
    python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-parity \
    > $O/r02_bench_under_ncu2.log 2>&1; echo "ncu_dram=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16 -s 1200 -c 6 \
    -o $O/r02_final_gemm_full python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-parity \
    > $O/r02_bench_under_ncu3.log 2>&1; echo "ncu_gemm_full=$?"
ls -la $O | tail -12
