T=r02zz; O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_unet_b16.json 2> $O/${T}_bench.err; tail -2 $O/${T}_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02zz_bench_unet_b16.json"))
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["clocks"], d["gpu_launches_per_step"], d["roofline"]["frac"], d["roofline_wgrad"]["frac"], d["cpu_baseline"]["value"])
print(d["diagnostics"])
P
