python -m pytest tests/test_reference_wrappers_gpu.py tests/test_plain_gpu.py -m gpu -q --timeout=900 2>&1 | tail -12
