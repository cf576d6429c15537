#!/bin/bash
# `ncu --set full` of representative launches (run under gpurun, one GPU), labelled by layer; the report comes back in
# gpurun_out/ and tools/summarize_ncu_full.py turns it into profiles/r2_*_ncu_full_summary.txt
set -u
mkdir -p gpurun_out
OPS=head.cls.0.2,backbone.dark4.1.m.0.cv1,backbone.dark5.1.m.0.cv1,backbone.dark3.1.m.0.cv2,backbone.dark2.1.m.0.cv1,neck.h4.cv1,backbone.dark5.2.m
timeout 300 python tools/profile_ops.py --ops $OPS > gpurun_out/plain_ops_full.log 2>&1 || exit 1
# (the warm pass launches ~61 matching kernels first: the window starts a little early, the LAST seven captured launches are
# the selected ops in plan order -- summarize_ncu_full.py aligns the labels from the end)
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_halo|conv_igemm|sppf_pool" -s 55 -c 14 -f -o gpurun_out/r2_full \
    python tools/profile_ops.py --ops $OPS > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
