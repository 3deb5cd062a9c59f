T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
timeout 900 $T bench.py --gpus 8 --steps 40 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -2 gpurun_out/r2_bench_n8.err
python -c "
import json
for l in open('gpurun_out/r2_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('hex200 n8', d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['ok'], d['config']['exchange'], d['config']['setup_s'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0})"
timeout 1500 $T bench.py --gpus 8 --steps 20 --warmup 3 --size 271 > gpurun_out/r2_bench_config5_n8.json 2> gpurun_out/r2_bench_config5_n8.err; tail -2 gpurun_out/r2_bench_config5_n8.err
python -c "
import json
for l in open('gpurun_out/r2_bench_config5_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print('config5 n8', d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['ok'], d['config']['exchange'], d['config']['setup_s'], {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0})"
