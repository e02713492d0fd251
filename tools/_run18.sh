cd /root/repo
timeout 300 python -m pytest tests/test_tracker_gpu.py -m gpu -x -q -k "max_report" 2>&1 | tail -30
