# Round-2 evidence run on one B200: options throughput, bench line, reference arm, ncu launch lists, ncu --set full of the
# dominant kernel (each ncu command only after the same command exited 0 without ncu).
set -x
python tests/gpu_options_throughput.py > gpurun_out/r2_options.log 2>&1; cat gpurun_out/r2_options.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
REPS=1 python tests/gpu_sweep.py configs > gpurun_out/r2_cfg_plain.log 2>&1 && \
REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_configs.csv python tests/gpu_sweep.py configs > gpurun_out/r2_ncu_configs.log 2>&1
python tests/gpu_sweep.py one markov 32 > gpurun_out/r2_one_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 1 -o gpurun_out/prof_radix_r2 -f \
    python tests/gpu_sweep.py one markov 32 > gpurun_out/r2_ncu_full.log 2>&1
python tests/gpu_sweep.py one markov 32 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_inv_walk2 -c 1 -o gpurun_out/prof_invwalk_r2 -f \
    python -m pytest tests/test_gpu_inverse.py -q -k "full_size and markov" > gpurun_out/r2_ncu_inv.log 2>&1
ls -la gpurun_out/*.ncu-rep
