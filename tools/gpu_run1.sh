cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2a_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
PANO_NO_TMA=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_c2_notma.json 2> gpurun_out/r2a_bench_c2_notma.err
PANO_FE_NO_WORDS=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_c2_nowords.json 2> gpurun_out/r2a_bench_c2_nowords.err
timeout 300 python bench.py --workload config1 --no-cpu-baseline > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
tail -3 gpurun_out/r2a_pytest.log
