# S4 (BASELINE configs[3]) weak-scaling point at N GPUs: bench.py --scene s4, 16 spp per GPU and step
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --scene s4 --spp 16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_s4_n1.json 2> gpurun_out/r2_s4_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --scene s4 --spp 16 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_s4_n$N.json 2> gpurun_out/r2_s4_n$N.err
fi
python -c "
import json; d=json.load(open('gpurun_out/r2_s4_n$N.json')); print('S4 N=$N', d['value']/1e6, 'M samples/s', d['ms_per_step'], 'ms/step', (d.get('per_rank') or {}).get('render_ms'))"
