#!/bin/bash
# correctness at 4 ranks + timers at 8/4 + plain LET at 8/4
timeout 300 python -m pytest tests/test_distributed.py -m gpu -x -q -k "domain" > gpurun_out/pytest_let4.log 2>&1; tail -3 gpurun_out/pytest_let4.log
bash tools/let_timers.sh 8 4
OUT=gpurun_out/letscale2 bash tools/let_scale.sh 8 2>&1 | grep -v "let rank"
