# Final round-2 validation on one B200: whole GPU suite, the bench line, the reference arm, the 9-bit digit variant's ncu capture.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -2 gpurun_out/r2f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err
BWTC_RADIX9=1 python tests/gpu_sweep.py one markov 32 > gpurun_out/r2f_one9_plain.log 2>&1 && \
BWTC_RADIX9=1 ncu --set full --clock-control none --import-source on -k regex:k_radix_pass -s 3 -c 1 -o gpurun_out/prof_radix9_r2 -f \
    python tests/gpu_sweep.py one markov 32 > gpurun_out/r2f_ncu_full9.log 2>&1
tail -3 gpurun_out/r2f_one9_plain.log
timeout 300 python tests/gpu_soak.py > gpurun_out/r2f_soak.log 2>&1; tail -3 gpurun_out/r2f_soak.log
ls -la gpurun_out/*.ncu-rep
