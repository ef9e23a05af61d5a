set -x
mkdir -p gpurun_out
python -m pytest tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_baseline_configs_gpu.py tests/test_multi_device_gpu.py -m gpu -q --maxfail=30 --timeout=900 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc $?"
tail -8 gpurun_out/pytest_r2d.log
python tools/sampler_sweep.py 1 2>&1 | grep -E "cluster=(auto|1|2|4) waves=auto chunks=(8|4)" > gpurun_out/sampler_sweep_r2d.log
cat gpurun_out/sampler_sweep_r2d.log
PQB200_LAP_CLUSTER=1 ncu --set full --clock-control none --import-source on -k regex:laplace -c 1 -o gpurun_out/ncu_sampler_c5_r2d -f python tools/profile_sampler.py 1000 > gpurun_out/ncu_sampler.log 2>&1
tail -3 gpurun_out/ncu_sampler.log
