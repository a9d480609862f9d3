#!/bin/bash
# Everything that needs the 8 GPUs of one box, in one gpurun call: scaling 1/2/4/8 of the default bench (BASELINE configs[1]),
# BASELINE config 4 (UNet 720x960, global batch 64 on 2/4/8 GPUs), config 5 (eval, 64 x 720x960 on 8 GPUs), the all-reduce
# micro-benchmark and the interleaved comparison of the exchange variants.   tools/multi_gpu_suite.sh r02l
TAG=${1:-mg}; OUT=gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520 "$@"; }
show() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        print(f"{f}: {d['value']:.1f} img/s, {d['ms_per_step']:.3f} ms/step, e2e {d['e2e']['value']:.1f}, n_gpus {d['n_gpus']}, "
              f"sm {d.get('clocks', {}).get('sm_mhz')}, dp_check {d.get('dp_check')}")
    except Exception as e:
        print(f, "FAILED", e)
PY
}
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_unet_b16_n1.json 2> $OUT/${TAG}.err; show $OUT/${TAG}_bench_unet_b16_n1.json
for n in 2 4 8; do
  tr $n bench.py --gpus $n --steps 20 --warmup 5 --no-kernel-timing > $OUT/${TAG}_bench_unet_b16_n$n.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_bench_unet_b16_n$n.json
done
CVB_DP_BACKEND=nccl tr 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-kernel-timing > $OUT/${TAG}_bench_unet_b16_n8_nccl.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_bench_unet_b16_n8_nccl.json
tr 8 tools/bench_allreduce.py 138 8 2>> $OUT/${TAG}.err | grep -E " MB " | tee $OUT/${TAG}_allreduce_8gpu.txt
tr 8 tools/dp_ab.py 6 8 2>> $OUT/${TAG}.err | grep -vE "^\*|OMP_NUM|^$|NCCL version" | tee $OUT/${TAG}_dp_ab_8gpu.txt | cut -c1-160
# BASELINE config 4: UNet training at 3x720x960, global batch 64
for nb in "8 8" "4 16" "2 32"; do set -- $nb
  tr $1 bench.py --gpus $1 --batch $2 --height 720 --width 960 --steps 8 --warmup 3 --no-kernel-timing > $OUT/${TAG}_bench_unet_720x960_n$1.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_bench_unet_720x960_n$1.json
done
# BASELINE config 5: eval, batch 64 x 3x720x960 over 8 GPUs, confusion-matrix mIoU
for m in unet segnet; do
  tr 8 bench.py --gpus 8 --mode eval --model $m --batch 8 --height 720 --width 960 --steps 10 --warmup 3 > $OUT/${TAG}_bench_${m}_eval_720x960_n8.json 2>> $OUT/${TAG}.err; show $OUT/${TAG}_bench_${m}_eval_720x960_n8.json
done
grep -iE "error|Traceback" $OUT/${TAG}.err | head -5
