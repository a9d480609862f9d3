T=r02w; O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; tail -3 $O/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -3 $O/${T}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 --graph > $O/${T}_bench_unet_b16.json 2> $O/${T}_bench.err; tail -2 $O/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err
timeout 200 python bench.py --model segnet --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_segnet_b16.json 2>> $O/${T}_bench.err
timeout 200 python bench.py --mode eval --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_unet_eval_b16.json 2>> $O/${T}_bench.err
timeout 200 python bench.py --mode eval --model segnet --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_segnet_eval_b16.json 2>> $O/${T}_bench.err
timeout 300 python bench.py --batch 8 --height 720 --width 960 --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench_unet_720x960_b8.json 2>> $O/${T}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $O/${T}_launches_unet_b16.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-kernel-timing > $O/${T}_ncu_launches.log 2>&1
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02w_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], round(d["value"],1), d["unit"], round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("clocks",{}).get("sm_mhz"), d.get("gpu_launches_per_step"), (d.get("graph_replay") or {}).get("value"))
    except Exception as e: print(f, "FAILED", e)
P
