#!/bin/bash
# last-minute sanity of the final binaries on a GPU: the e2e application through the component, one pool and three pools
mkdir -p gpurun_out
cd /tmp
PARSEC_MCA_device_b200_enabled=1 timeout 20 /root/repo/oracle/_ref/bin/ex05_b200 -K 512 -t 65536 -m gpu -c 8 -w -r 2 > /root/repo/gpurun_out/r02_sanity_p1.json 2>/root/repo/gpurun_out/r02_sanity_p1.err; echo "rc=$?"
PARSEC_MCA_device_b200_enabled=1 PARSEC_MCA_device_b200_nvtx=1 timeout 20 /root/repo/oracle/_ref/bin/ex05_b200 -K 256 -t 65536 -m gpu -c 8 -w -r 2 -P 3 > /root/repo/gpurun_out/r02_sanity_p3.json 2>/root/repo/gpurun_out/r02_sanity_p3.err; echo "rc=$?"
cut -c1-700 /root/repo/gpurun_out/r02_sanity_p1.json /root/repo/gpurun_out/r02_sanity_p3.json
