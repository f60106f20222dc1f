# multi-GPU check + weak-scaling bench on N GPUs of one box:  gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
set -x
N=${1:-2}
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/mgpu_check.py > gpurun_out/mgpu_${N}.log 2>&1; tail -6 gpurun_out/mgpu_${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; cut -c1-700 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 200 --warmup 5 --temperature 0 > gpurun_out/bench_${N}gpu_T0.json 2> gpurun_out/bench_${N}gpu_T0.err; cut -c1-700 gpurun_out/bench_${N}gpu_T0.json; tail -2 gpurun_out/bench_${N}gpu_T0.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/bench_1gpu_same_box.json 2>/dev/null; cut -c1-300 gpurun_out/bench_1gpu_same_box.json
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --temperature 0 > gpurun_out/bench_1gpu_same_box_T0.json 2>/dev/null; cut -c1-300 gpurun_out/bench_1gpu_same_box_T0.json
