# usage: benchcfg.sh tag cfg [env...]
tag=$1; c=$2; shift 2
env "$@" timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_cfg$c.log 2>&1
python - "$tag" "$c" "$*" <<'PY'
import json,sys
tag,c,e=sys.argv[1:4]
for l in open("gpurun_out/%s_cfg%s.log"%(tag,c)):
    if l.startswith("{"):
        d=json.loads(l); print(tag, "cfg",c, e, "value %.3f ms/step %.3f frac %.4f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"]))
PY
