exec > gpurun_out/run23.log 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python -m pytest tests/test_gpu_parity.py -q -x -s -k "torch_default_tf32" 2>&1 | grep "parameter-gradient\|passed\|failed"
python tools/profile_timeline.py --B 2 --H 480 --W 854 --nprod 2 --steps 10 2>&1 | grep "kernel time\|k_conv64\|k_stem\|k_pool"
python tools/profile_timeline.py --B 8 --H 96 --W 96 --resize --nprod 2 --steps 10 2>&1 | grep "kernel time"
