python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err
python bench.py --workload ns_sgpr --no-secondary > gpurun_out/r2z_bench_ns_sgpr.json 2> gpurun_out/r2z_bench_ns_sgpr.err
python bench.py --workload ns_sgpr --prec fp32 --no-secondary > gpurun_out/r2z_bench_ns_sgpr_fp32.json 2> gpurun_out/r2z_bench_ns_sgpr_fp32.err
python bench.py --workload cfg1_sgpr --no-secondary > gpurun_out/r2z_bench_cfg1_sgpr.json 2> gpurun_out/r2z_bench_cfg1_sgpr.err
python bench.py --prec fp32 --no-secondary > gpurun_out/r2z_bench_fp32.json 2> gpurun_out/r2z_bench_fp32.err
