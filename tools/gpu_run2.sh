cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests/test_gpu_fullsize_cv2.py -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/r2b_pytest_fullsize.log
timeout 600 python bench.py > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
tail -5 gpurun_out/r2b_pytest_fullsize.log; tail -3 gpurun_out/r2b_bench_c2.err; head -c 600 gpurun_out/r2b_bench_c2.json
