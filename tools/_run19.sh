cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_length_oracle" 2>&1 | tail -6
timeout 300 python tools/sweep_shapes.py --out gpurun_out/sweep19.json 1:8:128/c@2368 --shape shelf 2>&1 | grep mfps | cut -c1-120
timeout 300 python tools/sweep_shapes.py --out gpurun_out/sweep19c.json --shape campus --frames 2000 1:8:128/c@2368 2>&1 | grep mfps | cut -c1-120
