set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 scripts/mgpu_check.py > gpurun_out/mgpu23.log 2>&1; tail -8 gpurun_out/mgpu23.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench23_2gpu.json 2> gpurun_out/bench23_2gpu.err; cat gpurun_out/bench23_2gpu.json | cut -c1-1200; tail -3 gpurun_out/bench23_2gpu.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/bench23_1gpu.json 2>/dev/null; cut -c1-400 gpurun_out/bench23_1gpu.json
