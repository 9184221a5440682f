# Final round-2 validation and evidence on one B200: smoke(), whole GPU suite, bench line, reference arm, ncu launch lists of
# the five BASELINE configs, ncu --set full of the dominant kernel (each ncu command only after the same command exited 0
# without ncu), soak.
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -2 gpurun_out/r2f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err
if [ -z "$SKIP_NCU" ]; then
REPS=1 python tests/gpu_sweep.py configs > gpurun_out/r2f_cfg_plain.log 2>&1 && \
REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2f_launches_configs.csv python tests/gpu_sweep.py configs > gpurun_out/r2f_ncu_configs.log 2>&1
cat gpurun_out/r2f_cfg_plain.log
python tests/gpu_sweep.py one markov 32 > gpurun_out/r2f_one_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 1 -o gpurun_out/prof_radix_r2f -f \
    python tests/gpu_sweep.py one markov 32 > gpurun_out/r2f_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
timeout 300 python tests/gpu_soak.py > gpurun_out/r2f_soak.log 2>&1; tail -3 gpurun_out/r2f_soak.log
