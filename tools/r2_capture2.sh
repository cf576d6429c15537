#!/bin/bash
# Round-2 second capture: role-wait profile of representative layers + per-launch ncu counters of one eager pass (all ops)
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout 300 python tools/conv_profile.py --filter dark3.1.m.0.cv,dark4.1.m.0.cv,dark5.1.m.0.cv,dark4.1.cv1,dark5.1.cv2,neck.h6.cv1,head.cls.1.2,head.boxcls.1.0,dark2.1.m.0.cv,neck.h1.cv1,dark3.1.cv1 > gpurun_out/role_waits_$TAG.log 2>&1; echo "profile rc=$?"
timeout 300 python tools/profile_ops.py --ops all > gpurun_out/plain_ops_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.max,sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg,lts__t_sectors_srcunit_tex_op_read.sum \
    --clock-control none -s 61 -c 80 --csv --log-file gpurun_out/percounter_$TAG.csv python tools/profile_ops.py --ops all > gpurun_out/ncu_percounter_$TAG.log 2>&1
echo "ncu rc=$?"
