# usage (through gpurun --gpus N): bash scripts/r2_scale.sh N TAG [extra bench flags...]  -> gpurun_out/TAG_n{N}_{weak,strong}.json
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
for mode in weak strong; do
  timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29511 + RANDOM % 400)) \
    bench.py --gpus $N --steps 40 --warmup 5 --scaling $mode "$@" > gpurun_out/${TAG}_n${N}_${mode}.json 2> gpurun_out/${TAG}_n${N}_${mode}.err
  echo "$mode rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/${TAG}_n${N}_${mode}.json") if l.startswith("{")][-1]
    print("${mode}", d["n_gpus"], "value %.1f ms %.3f e2e %.1f opt %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_optimizer"]["value"]), {k: round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("${mode} failed", e)
PY
done
