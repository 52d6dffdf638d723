cd $GRAFT_REPO_ROOT
T=r3l
K="golden or library_weights or feather or band_counts or odd_geometry or pageable or mask_update or batched or gain_map or front_end_golden or ring_epilogue or fit2final or yuyv or reads_only"
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" 2>&1 | tail -3)
(timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$K" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${T}_memcheck.log)
grep -E "ERROR SUMMARY|memcheck exit|passed|failed|Invalid|out of bounds" gpurun_out/${T}_memcheck.log | head -20
