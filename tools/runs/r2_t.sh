#!/bin/bash
# round 2, step t: lidar block of PoseOptimization on the device
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pose_opt.py -x -q > gpurun_out/r2_t_poseopt.log 2>&1; tail -30 gpurun_out/r2_t_poseopt.log
