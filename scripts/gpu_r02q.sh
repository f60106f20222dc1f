set -x
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -x -k "slab or multi_gpu" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 scripts/mgpu_check.py > gpurun_out/r02q_mgpu_check_${N}gpu.log 2>&1; grep "mgpu_check\|MGPU" gpurun_out/r02q_mgpu_check_${N}gpu.log; tail -3 gpurun_out/r02q_mgpu_check_${N}gpu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 scripts/mgpu_trace.py 2>&1 | grep "rank" | tee gpurun_out/r02q_mgpu_trace_${N}gpu_bulk.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29742 scripts/mgpu_trace.py '{"halo_bulk":0}' 2>&1 | grep "rank" | tee gpurun_out/r02q_mgpu_trace_${N}gpu_nobulk.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29743 scripts/mgpu_trace.py '{"face_after":0}' 2>&1 | grep "rank" | tee gpurun_out/r02q_mgpu_trace_${N}gpu_bulk_face_first.log
