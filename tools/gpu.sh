#!/bin/bash
# build first (a stale .so travels silently otherwise), then run the command on a B200 box:  tools/gpu.sh [--gpus N] TIMEOUT 'command'
set -e
cd "$(dirname "$0")/.."
python build.py >/dev/null
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
exec /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
