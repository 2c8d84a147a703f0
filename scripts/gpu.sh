#!/bin/bash
# usage: scripts/gpu.sh [--timeout S] [--gpus N] -- '<command>'   (rebuilds libdbmm.so first: the built .so ships with the snapshot)
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
