#!/bin/bash
B=oracle/_ref/bin/ex05_b200
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['cores'], d['times_s'], d['errors'], {k: d['b200'][k] for k in ('kernel_launches','peer_pulls','manager_entries')})"; }
echo "== 2 dev c=16"; PARSEC_MCA_device_b200_enabled=2 timeout 100 $B -K 8192 -t 65536 -m gpu -c 16 -r 6 2>&1 | tail -1 | show
echo "== 2 dev c=32"; PARSEC_MCA_device_b200_enabled=2 timeout 100 $B -K 8192 -t 65536 -m gpu -c 32 -r 6 2>&1 | tail -1 | show
echo "== 2 dev c=16 idle 20ms"; PARSEC_MCA_device_b200_idle_us=20000 PARSEC_MCA_device_b200_enabled=2 timeout 100 $B -K 8192 -t 65536 -m gpu -c 16 -r 6 2>&1 | tail -1 | show
echo "== 2 dev c=16 serial completion"; PARSEC_MCA_device_b200_parallel_completion=0 PARSEC_MCA_device_b200_enabled=2 timeout 100 $B -K 8192 -t 65536 -m gpu -c 16 -r 6 2>&1 | tail -1 | show
