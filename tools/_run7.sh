timeout 300 python -m pytest tests/test_gpu_ssm_kalman.py -m gpu -q -x -p no:cacheprovider -k "time_sharded" 2>&1 | tail -15
