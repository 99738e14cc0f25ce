# Full-size parity tests (C4 250k atoms, C2 64 x 16000, C5-sized shard) and the whole GPU suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -20
