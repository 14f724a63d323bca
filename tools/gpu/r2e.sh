mkdir -p gpurun_out
ABR_SYMMETRIC=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2e_launches_sym.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log | cut -c1-300
(timeout 700 python -m pytest tests/test_gpu_symmetric.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -25) > gpurun_out/r2e_tests.log
tail -25 gpurun_out/r2e_tests.log
