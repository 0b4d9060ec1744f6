set -x
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --scaling strong --no-cpu-baseline --no-retrieval > gpurun_out/r02_strong_n$N.json 2> gpurun_out/r02_strong_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 141 --no-cpu-baseline > gpurun_out/r02_config5_n8.json 2> gpurun_out/r02_config5_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline --no-retrieval > gpurun_out/r02_weak_n8.json 2> gpurun_out/r02_weak_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload videomae --no-cpu-baseline --no-retrieval > gpurun_out/r02_videomae_n8.json 2> gpurun_out/r02_videomae_n8.err
tail -c 400 gpurun_out/r02_strong_n8.json; tail -c 300 gpurun_out/r02_strong_n8.err
