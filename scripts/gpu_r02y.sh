set -x
N=${1:-4}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/mgpu_check.py > gpurun_out/r02y_mgpu_check_${N}gpu.log 2>&1; grep "mgpu_check\|MGPU" gpurun_out/r02y_mgpu_check_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02y_bench_${N}gpu.json 2> gpurun_out/r02y_bench_${N}gpu.err; cut -c1-300 gpurun_out/r02y_bench_${N}gpu.json; tail -2 gpurun_out/r02y_bench_${N}gpu.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-extra > gpurun_out/r02y_bench_1gpu_same_box_as_${N}.json 2>/dev/null; cut -c1-300 gpurun_out/r02y_bench_1gpu_same_box_as_${N}.json
