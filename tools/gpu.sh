#!/bin/bash
# Rebuild everything that travels to the GPU box (product library, oracle libraries incl. the timing builds), then run a command there.
# usage: tools/gpu.sh [--gpus N] [--timeout S] '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
exec /usr/local/graft/bin/gpurun "$@"
