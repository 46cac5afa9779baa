mkdir -p gpurun_out
run() {  # tag, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 $2 --no-cpu-baseline > gpurun_out/r2n8_$1.json 2> gpurun_out/r2n8_$1.err
  python -c "
import json; d=json.load(open('gpurun_out/r2n8_$1.json')); pr=d['per_rank']; print('$1', 'value',round(d['value']/1e6,1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']/1e6,1) if d['e2e'] else None, 'imbalance',round(pr['imbalance'],4),'calls',pr['render_calls'],'render_ms',[round(x) for x in pr['render_ms']],'closest_ms',[round(x) for x in pr['trace_closest_ms']])" || tail -5 gpurun_out/r2n8_$1.err
}
run s3_static "--steps 8 --warmup 3 --tiles static"
run s3_dynamic "--steps 8 --warmup 3 --tiles dynamic"
run s4_static "--scene s4 --spp 16 --steps 3 --warmup 1 --tiles static --no-e2e"
run s4_dynamic "--scene s4 --spp 16 --steps 3 --warmup 1 --tiles dynamic --no-e2e"
