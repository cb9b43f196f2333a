#!/bin/bash
# dev helper: rebuild (no-op when up to date), then run a command on the GPU box.  usage: tools/gpu.sh <timeout> '<cmd>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -iE "error|^\+ " || true
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
