set -x
mkdir -p gpurun_out
python -m pytest tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q --maxfail=30 --timeout=900 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc $?"
tail -8 gpurun_out/pytest_r2e.log
python tools/sampler_sweep.py 1 > gpurun_out/sampler_sweep_r2e.log 2>&1
cat gpurun_out/sampler_sweep_r2e.log
