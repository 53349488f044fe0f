mkdir -p gpurun_out
run() { tag=$1; shift; (timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 10 --warmup 3 --replay-blocks-multi 0 --replay-blocks 0 > gpurun_out/r2_bench_n8_$tag.json 2> gpurun_out/r2_bench_n8_$tag.err); python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench_n8_$tag.json')); c=d['clocks']
print('$tag', 'value', round(d['value']/1e6,2), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,2))
print(' kernel', c.get('kernel_ms_per_rank')); print(' step  ', c.get('step_ms_per_rank'))
for t in c.get('timeline_per_rank') or []: print('  ', t)
"; }
PORT=29520 run sampler A=1
PORT=29521 run nosampler KGV_BENCH_NO_SAMPLER=1
