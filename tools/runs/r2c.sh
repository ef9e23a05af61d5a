set -x
mkdir -p gpurun_out
python tools/sampler_sweep.py 1 > gpurun_out/sampler_sweep_r2c.log 2>&1
python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc $?"
tail -15 gpurun_out/pytest_r2c.log
PQB200_LAP_CLUSTER=1 ncu --set full --clock-control none --import-source on -k regex:laplace -c 1 -o gpurun_out/ncu_sampler_c5_r2c -f python tools/profile_sampler.py 1000 > gpurun_out/ncu_sampler.log 2>&1
tail -3 gpurun_out/ncu_sampler.log
