timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider -k "loss" 2>&1 | tail -6
timeout -s KILL 200 python tools/loss_probe.py 2>&1 | tail -3
