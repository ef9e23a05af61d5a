set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/pytest_r2j.log
python tools/sampler_sweep.py 1 125,1000,3000,20000 2>&1 | grep -E "default|TAIL': '0'" | tee gpurun_out/sampler_sweep_r2j.log
python tools/rep_cases.py 2>&1 | tee gpurun_out/rep_cases_r2j.log
