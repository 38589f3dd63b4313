exec > gpurun_out/run2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^k_loss$|^k_bwd$" -c 4 -o gpurun_out/r02_affine_core2 python tools/profile_step.py --D 2 --steps 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
