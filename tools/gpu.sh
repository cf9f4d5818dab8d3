#!/bin/bash
# build the library (if stale) and run a command on the GPU box:  tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
