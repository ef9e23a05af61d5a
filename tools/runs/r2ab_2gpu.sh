# 2-GPU box, final sampler host path (rng_replay.h, device-side order of the second launch)
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_device_gpu.py tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_nccl_gpu.py -m gpu -q --timeout=600 > gpurun_out/pytest_r2ab.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_r2ab.log
SWEEP_REPS=5 python tools/sampler_sweep.py 1,2 1000,20000 '[{}]' 2>&1 | tee gpurun_out/sampler_2gpu_r2ab.log
