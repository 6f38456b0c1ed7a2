# slab mode over NCCL on N GPUs (N = number of visible GPUs): parity vs one handle, then the timed run
cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c5 --points 10000000 --steps 2 --warmup 1 --verify 2> gpurun_out/r12_c5_10m_n$N.err | tee gpurun_out/r12_c5_10m_n$N.json
tail -3 gpurun_out/r12_c5_10m_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --points 50000000 --steps 3 --warmup 1 2> gpurun_out/r12_c5_50m_n$N.err | tee gpurun_out/r12_c5_50m_n$N.json
tail -3 gpurun_out/r12_c5_50m_n$N.err
