set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -15
timeout 300 python scripts/mt_sweep.py --n 256 --temps 50,150,250,330,400,500 --equil 3000 --meas 3000 --every 20 > gpurun_out/r02p_mt_sweep_c3.log 2>&1; cat gpurun_out/r02p_mt_sweep_c3.log | cut -c1-400
