timeout 500 python -m pytest tests/test_gpu_configs.py -q -k "two_gpus and npt" 2>&1 | tail -30
