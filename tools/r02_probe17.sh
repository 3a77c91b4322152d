#!/bin/bash
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/debug_rtt.py 2>&1 | grep -v "OMP\|\*\*\*\*\|Setting" | cut -c1-700
