exec > gpurun_out/run3.log 2>&1
python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
echo "== 3 CTAs/SM"; python tools/profile_timeline.py --B 2 --H 480 --W 854 --nprod 2 --steps 10 2>&1 | grep "k_pool\|k_stem\|kernel time"
sed -i 's/__launch_bounds__(RCF_BLOCK, 3) k_pool_bwd_nhwc/__launch_bounds__(RCF_BLOCK, 2) k_pool_bwd_nhwc/' rcf_unsupvideoseg_b200/csrc/rcf_pool.cu
python -m rcf_unsupvideoseg_b200.build 2>&1 | tail -1
echo "== 2 CTAs/SM"; python tools/profile_timeline.py --B 2 --H 480 --W 854 --nprod 2 --steps 10 2>&1 | grep "k_pool\|k_stem\|kernel time"
