set -x
python -m pytest tests/test_ops_gpu.py -x -q -k "pool or bilinear or last_block" > gpurun_out/r02o_pytest_ops.log 2>&1; tail -3 gpurun_out/r02o_pytest_ops.log
for s in "16 180 240 128" "16 90 120 256" "16 45 60 512" "16 360 480 64"; do python tools/bench_bn.py $s 2>&1 | grep -E "shape|pool_bwd|bilinear"; done > gpurun_out/r02o_ew.txt 2>&1
cat gpurun_out/r02o_ew.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02o_bench_unet_b16.json 2> gpurun_out/r02o_bench.err; tail -2 gpurun_out/r02o_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02o_bench_unet_b16.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d.get("graph_replay",{}).get("value"))
for k,v in d["kernels"].items(): print(k, v["avg_us"], v["launches_per_step"], v["frac"], v["share_of_step"])
P
