# slab mode over NCCL on all visible GPUs: parity on 10 M points, timed run on 50 M points
cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --workload c5 --points 10000000 --steps 1 --warmup 1 --verify 2> gpurun_out/r17_c5_10m_n$N.err | tee gpurun_out/r17_c5_10m_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('10M n=%d'%d['n_gpus'], 'verified', d['verified'] and d['verified']['identical_to_single_handle_on_every_rank'], 'ms', round(d['ms_per_step'],1), d['stage_ms_max_over_ranks'])
"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --workload c5 --points 50000000 --steps 3 --warmup 1 2> gpurun_out/r17_c5_50m_n$N.err | tee gpurun_out/r17_c5_50m_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('50M n=%d'%d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d['stage_ms_max_over_ranks'], 'bytes', d['exchanged_bytes_per_step_rank0'])
"
tail -2 gpurun_out/r17_c5_50m_n$N.err
