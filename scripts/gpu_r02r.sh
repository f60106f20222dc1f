set -x
N=${1:-2}
for O in '{"face_after":2}' '{"face_after":3}' '{"face_after":5}'; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 scripts/mgpu_trace.py "$O" 2>&1 | grep "rank 0" | tee -a gpurun_out/r02r_mgpu_trace_${N}gpu.log
done
