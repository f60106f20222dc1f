# multi-GPU check + weak / strong scaling bench on N GPUs of one box:  gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [tag]'
set -x
N=${1:-2}
TAG=${2:-r02y}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/mgpu_check.py > gpurun_out/${TAG}_mgpu_check_${N}gpu.log 2>&1; grep "mgpu_check\|MGPU" gpurun_out/${TAG}_mgpu_check_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; cut -c1-300 gpurun_out/${TAG}_bench_${N}gpu.json; tail -2 gpurun_out/${TAG}_bench_${N}gpu.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-extra > gpurun_out/${TAG}_bench_1gpu_same_box_as_${N}.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_1gpu_same_box_as_${N}.json
# per-item traces of every rank (where the multi-GPU step loses time)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 scripts/mgpu_trace.py 2>&1 | grep "rank" | tee gpurun_out/${TAG}_mgpu_trace_${N}gpu.log
