# after the order of a wave's second launch moved to the device (lap_rank_kernel) and tail 3 -> 4
set -x
mkdir -p gpurun_out
python -m pytest tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_baseline_configs_gpu.py tests/test_reference_wrappers_gpu.py tests/test_multi_device_gpu.py -m gpu -q --timeout=600 > gpurun_out/pytest_r2t.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_r2t.log
SWEEP_REPS=5 python tools/sampler_sweep.py 1 125,444,1000,3000,20000 '[{}, {"PQB200_LAP_TAIL": "0"}]' 2>&1 | tee gpurun_out/sampler_sweep_r2t.log
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sampler_sweep.py 1 600 '[{}]' 2>&1 | tail -4
