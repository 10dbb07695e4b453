mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu -k "metis or match_single_domain[fused]" > gpurun_out/t_mgpu.log 2>&1; echo "tests exit $?" >> gpurun_out/t_mgpu.log
grep "^E   \|^FAILED\|passed\|failed\|skipped" gpurun_out/t_mgpu.log | head
