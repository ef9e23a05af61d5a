python -m pytest tests/test_multi_device_gpu.py tests/test_nccl_gpu.py tests/test_sampling_gpu.py -m gpu -q --timeout=900 2>&1 | tail -4
python tools/sampler_sweep.py 1,2 1000,20000 2>&1 | grep -E "default"
