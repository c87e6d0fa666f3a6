set -x
timeout 600 python -m pytest tests/test_comm_gpu.py -x -q -s 2>&1 | tail -15
for coll in fused nccl torch; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --collective $coll > gpurun_out/bench_n2_$coll.json 2> gpurun_out/bench_n2_$coll.err; echo rc=$?; tail -c 600 gpurun_out/bench_n2_$coll.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n2_$coll.json").read().strip().splitlines()[-1])
    print("$coll", d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("collective"), d["parity"])
except Exception as e: print("ERR", e)
PY
done
