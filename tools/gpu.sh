#!/bin/bash
# build, then run a command on the GPU box: tools/gpu.sh [timeout_s] -- 'command'
set -e
cd /root/repo
make -C horde-ad_b200/csrc -j8 2>&1 | grep -i "error" -A3 && exit 1
T=600
if [ "$1" != "--" ]; then T=$1; shift; fi
shift
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
