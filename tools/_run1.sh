mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_all.log 2>&1; echo "tests exit $?" >> gpurun_out/t_all.log
tail -6 gpurun_out/t_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-200
