#!/bin/bash
# One 8-GPU gpurun call: hardware verification of the sharded path on 8 and 4 GPUs, then the
# driver-style bench at N = 8 and N = 4 (N = 2 is run on 2-GPU boxes, which are cheaper).
set -u
export TAG=${TAG:-r02_n8}
o=gpurun_out
CHECK_MODES="pull pull2 push nccl" NOBENCH=1 tools/multi_gpu_campaign.sh 8
CHECK_MODES="pull" NOBENCH=1 tools/multi_gpu_campaign.sh 4
CHECK_MODES="" tools/multi_gpu_campaign.sh 8 4
