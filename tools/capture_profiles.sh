#!/bin/bash
# Final round-1 evidence run on one B200 (under gpurun): tests, smoke, bench, ncu launch list and
# full captures of the dominant kernels.  Outputs land in gpurun_out/ and are copied to profiles/.
set -u
export EDB_NATIVE_LN=${EDB_NATIVE_LN:-1}
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu > $O/final_pytest_gpu.log 2>&1; echo "pytest_exit=$?" >> $O/final_pytest_gpu.log
tail -4 $O/final_pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
timeout 600 python bench.py --cpu-sample-seqs 1 > $O/final_bench.log 2>&1; echo "bench_exit=$?" >> $O/final_bench.log
tail -2 $O/final_bench.log | cut -c1-3000
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file $O/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline \
    > $O/final_bench_under_ncu.log 2>&1; echo "ncu_list=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_bf16 -s 700 -c 4 \
    -o $O/final_prof_gemm python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline \
    > $O/final_bench_under_ncu2.log 2>&1; echo "ncu_gemm=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ln_ -s 60 -c 4 \
    -o $O/final_prof_ln python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline \
    > $O/final_bench_under_ncu3.log 2>&1; echo "ncu_ln=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_ce_|k_sgd_momentum|k_splitk_reduce" -s 2 -c 5 \
    -o $O/final_prof_loss_optim python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline \
    > $O/final_bench_under_ncu4.log 2>&1; echo "ncu_loss_optim=$?"
ls -la $O | tail -14
