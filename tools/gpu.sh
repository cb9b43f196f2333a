#!/bin/bash
# dev helper: rebuild (no-op when up to date), then run a command on the GPU box.  usage: tools/gpu.sh <timeout> '<cmd>'
cd "$(dirname "$0")/.."
if ! python -c "import __graft_entry__ as g; g.build()" > /tmp/gvib200_build.log 2>&1; then
    grep -E "error" -A4 /tmp/gvib200_build.log | head -40
    echo "BUILD FAILED"
    exit 1
fi
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
