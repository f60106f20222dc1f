# strong scaling of BASELINE config 5 (sc 512^3) on N GPUs of one box:  gpurun --gpus N -- 'bash scripts/gpu_multi_strong.sh N [T0]'
set -x
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --strong --steps 50 --warmup 3 > gpurun_out/bench_strong_${N}gpu.json 2> gpurun_out/bench_strong_${N}gpu.err; cut -c1-500 gpurun_out/bench_strong_${N}gpu.json; tail -2 gpurun_out/bench_strong_${N}gpu.err
if [ "$2" = "T0" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus $N --strong --steps 50 --warmup 3 --temperature 0 > gpurun_out/bench_strong_${N}gpu_T0.json 2> gpurun_out/bench_strong_${N}gpu_T0.err; cut -c1-500 gpurun_out/bench_strong_${N}gpu_T0.json; tail -2 gpurun_out/bench_strong_${N}gpu_T0.err
fi
