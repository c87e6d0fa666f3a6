# usage: bash tools/n8_collective.sh [NGPUS]   -- config 2 with the fused all-reduce, then a small workload where the collective shows
N=${1:-8}
run() {  # workload collective steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $3 --warmup 3 --workload $1 --collective $2 > gpurun_out/bench_n${N}_$1_$2.json 2> gpurun_out/bench_n${N}_$1_$2.err
  echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$1_$2.json").read().strip().splitlines()[-1])
    print("$1 $2", "ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["config"].get("collective"), d.get("parity", {}).get("max_rel_err"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_n${N}_$1_$2.err").read()[-1500:])
PY
}
run c2 fused 20
run c2_small fused 50
run c2_small nccl 50
run c2_small torch 50
