python -m pytest tests/test_gpu_connect.py -q -k "per_ply" 2>&1 | tail -3
python tools/time_traj_grids.py 2>&1 | tail -3
