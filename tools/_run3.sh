exec > gpurun_out/run3.log 2>&1
python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python tools/profile_timeline.py --B 2 --H 480 --W 854 --nprod 2 --steps 10 2>&1 | grep -v Warn | head -12
python tools/profile_timeline.py --B 2 --H 480 --W 854 --nprod 3 --steps 10 2>&1 | grep -v Warn | head -10
