mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_driver_parity.py -q -m gpu -k "plot_file" > gpurun_out/t_drv.log 2>&1; echo "tests exit $?" >> gpurun_out/t_drv.log
grep "^E   \|^FAILED\|passed\|failed" gpurun_out/t_drv.log | head -30
