set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r2m_8gpu.json 2> gpurun_out/bench_r2m_8gpu.err; echo "bench rc $?"
tail -c 600 gpurun_out/bench_r2m_8gpu.err
python tools/sampler_sweep.py 1,2,4,8 1000,20000 2>&1 | grep -E "default" | tee gpurun_out/sampler_sweep_r2m_8gpu.log
python - <<'PY' 2>&1 | tee gpurun_out/inproc_8gpu_r2m.log
import time, numpy as np
from scipy.stats import unitary_group
import piquassoboost_b200 as pqb
A = np.ascontiguousarray(unitary_group.rvs(40, random_state=40))
small = np.ascontiguousarray(unitary_group.rvs(14, random_state=14))
pqb.GlynnPermanent(small).calculate(); print("n=14 automatic:", pqb.last_stats())
for r in range(3):
    t = time.perf_counter(); v = pqb.GlynnPermanent(A).calculate(); dt = time.perf_counter() - t
    print(f"n=40 in-process automatic ndev: {v} in {dt:.3f} s", pqb.last_stats())
B = np.ascontiguousarray(unitary_group.rvs(30, random_state=30))
for r in range(2):
    t = time.perf_counter(); v = pqb.GlynnPermanent(B).calculate(); dt = time.perf_counter() - t
    print(f"n=30 automatic: {dt*1e3:.3f} ms", pqb.last_stats())
PY
python -m pytest tests/test_multi_device_gpu.py tests/test_nccl_gpu.py -m gpu -q --timeout=600 2>&1 | tail -4
