#!/bin/bash
# Data-parallel limiter experiments on N GPUs of one box: same bench, one knob at a time.
#   tools/dp_experiments.sh 2 r02f      -> gpurun_out/r02f_dp2_<tag>.json
N=${1:-2}; TAG=${2:-dp}; STEPS=${3:-20}
run() {  # name, env assignments...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps $STEPS --warmup 5 --no-kernel-timing > gpurun_out/${TAG}_dp${N}_${name}.json 2>> gpurun_out/${TAG}_dp${N}.err
  python - "$name" gpurun_out/${TAG}_dp${N}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:18s} value {d['value']:8.1f} img/s  {d['ms_per_step']:7.3f} ms/step  e2e {d['e2e']['ms_per_step']:7.3f} ms  "
          f"sm {d['clocks']['sm_mhz']}  dp_check {d.get('dp_check', {}).get('grad_rel_err')} spread {d.get('dp_check', {}).get('param_spread_over_ranks_max_abs')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run no_exchange CVB_DP_PAYLOAD=none
run nccl CVB_DP_BACKEND=nccl
run nccl_one CVB_DP_BACKEND=nccl CVB_BUCKET_MB=100000
run nvlink_32 CVB_DP_BACKEND=nvlink CVB_ALLREDUCE_CTAS=32
run nvlink_64 CVB_DP_BACKEND=nvlink CVB_ALLREDUCE_CTAS=64
run nvlink_16 CVB_DP_BACKEND=nvlink CVB_ALLREDUCE_CTAS=16
run nvlink_one_148 CVB_DP_BACKEND=nvlink CVB_BUCKET_MB=100000 CVB_ALLREDUCE_CTAS=148
run nvlink_b8_64 CVB_DP_BACKEND=nvlink CVB_BUCKET_MB=8 CVB_ALLREDUCE_CTAS=64
run no_exchange_b CVB_DP_PAYLOAD=none
run nccl_b CVB_DP_BACKEND=nccl
run nvlink_64_b CVB_DP_BACKEND=nvlink CVB_ALLREDUCE_CTAS=64
run nvlink_one_148_b CVB_DP_BACKEND=nvlink CVB_BUCKET_MB=100000 CVB_ALLREDUCE_CTAS=148
