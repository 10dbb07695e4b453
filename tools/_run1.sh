mkdir -p gpurun_out
for g in 3 4 6 7 8; do
  HF_FUSED_CFG_G=$g timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/g_$g.log 2>&1
  echo "cfg_g $g: $(tail -1 gpurun_out/g_$g.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["ms_per_step"]/4,3), round(d["roofline"]["launch_ms"],3))' 2>&1 | tail -1)"
done
