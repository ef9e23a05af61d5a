python -m pytest tests/test_sampling_gpu.py tests/test_lossy_gpu.py tests/test_baseline_configs_gpu.py::test_c5_twenty_photons_forty_modes -m gpu -q --timeout=900 2>&1 | tail -5
python tools/sampler_sweep.py 1 1000,3000 2>&1 | grep -E "default"
