# after the host replay of the random stream moved to csrc/rng_replay.h: the sampler tests (shots identical to the
# reference), then call / kernel time of C5 under the launch-policy knobs
set -x
mkdir -p gpurun_out
python -m pytest tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_baseline_configs_gpu.py tests/test_reference_wrappers_gpu.py tests/test_multi_device_gpu.py -m gpu -q --timeout=600 > gpurun_out/pytest_r2s.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_r2s.log
SWEEP_REPS=5 python tools/sampler_sweep.py 1 1000,20000 '[{}, {"PQB200_LAP_TAIL": "3"}, {"PQB200_LAP_TAIL": "4"}, {"PQB200_LAP_TAIL": "5"}, {"PQB200_LAP_CHUNKS": "2"}, {"PQB200_LAP_CHUNKS": "8"}, {"PQB200_LAP_TAIL": "4", "PQB200_LAP_CHUNKS": "2"}]' 2>&1 | tee gpurun_out/sampler_sweep_r2s.log
SWEEP_REPS=5 python tools/sampler_sweep.py 1 125,444,3000 '[{}, {"PQB200_LAP_TAIL": "4"}]' 2>&1 | tee -a gpurun_out/sampler_sweep_r2s.log
