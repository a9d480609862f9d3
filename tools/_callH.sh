timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02t_pytest.log 2>&1; tail -4 gpurun_out/r02t_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02t_bench_unet_b16.json 2> gpurun_out/r02t_bench.err; tail -2 gpurun_out/r02t_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02t_bench_unet_b16.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d.get("graph_replay"))
for k,v in d["kernels"].items(): print(f"{k:28s} {v['avg_us']:8.1f} x{v['launches_per_step']:3d} frac {v['frac']:.3f} share {v['share_of_step']:.4f}")
print(d["diagnostics"], d["gpu_launches_per_step"])
P
