#!/bin/bash
# builds everything in-tree, then runs a command on the B200 box:  ./tools_gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
