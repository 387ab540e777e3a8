# usage (through gpurun --gpus N): bash scripts/r2_strong8.sh N TAG -> strong scaling with and without --overlap
N=$1; TAG=$2
mkdir -p gpurun_out
i=0
for extra in "" "--overlap"; do
  i=$((i + 1))
  name=${TAG}_n${N}_strong$( [ -n "$extra" ] && echo _overlap )
  timeout -k 10 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29611 + 37 * i)) \
    bench.py --gpus $N --steps 40 --warmup 5 --scaling strong $extra > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"
  python - gpurun_out/$name.json <<'PY'
import json, sys
try:
    d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
    print("value %.1f ms %.3f e2e %.1f opt %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["with_optimizer"]["value"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
except Exception as e:
    print("failed", e)
PY
done
