exec > gpurun_out/run13.log 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python tools/profile_host.py 2>&1 | grep -v Warn | head -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/profile_timeline.py --B 8 --H 96 --W 96 --resize --nprod 2 --steps 10 2>&1 | grep "kernel time"
