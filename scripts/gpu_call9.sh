set -x
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 scripts/mgpu_check.py > gpurun_out/mgpu9.log 2>&1; tail -20 gpurun_out/mgpu9.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench9_n2.json 2> gpurun_out/bench9_n2.err; cat gpurun_out/bench9_n2.json; tail -5 gpurun_out/bench9_n2.err
timeout 900 python scripts/quick_bench.py 256 100 > gpurun_out/quick_bench9.log 2>&1; cat gpurun_out/quick_bench9.log
