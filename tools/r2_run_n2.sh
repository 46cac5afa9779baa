mkdir -p gpurun_out
for t in static dynamic; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 2 --tiles $t --no-cpu-baseline > gpurun_out/r2n2_$t.json 2> gpurun_out/r2n2_$t.err
  python -c "
import json; d=json.load(open('gpurun_out/r2n2_$t.json')); print('$t', 'value',d['value']/1e6,'ms',d['ms_per_step'],'e2e',d['e2e']['value']/1e6 if d['e2e'] else None, d['per_rank'])" || tail -5 gpurun_out/r2n2_$t.err
done
