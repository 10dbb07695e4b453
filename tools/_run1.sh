mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_driver_parity.py -q -m gpu > gpurun_out/t_drv.log 2>&1; echo "tests exit $?" >> gpurun_out/t_drv.log
grep "^E   Assert\|^FAILED\|passed\|failed" gpurun_out/t_drv.log | head -20
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_new.log 2>&1; tail -1 gpurun_out/b_new.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["ms_per_step"]/4,3), round(d["roofline"]["launch_ms"],3))'
