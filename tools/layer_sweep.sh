#!/bin/bash
# Isolated layer timings under environment overrides (run under gpurun).  LAYERS = comma list of layer indices
# (tools/run_layer.py), every further argument line below is one configuration: `run VAR=value ...`.
# FVY_DBG=1 adds the cycle counters / %globaltimer timeline of every layer.
L=${LAYERS:-1,3,6,10,11,27,28,45,58}
run() { echo "== $*"; env "$@" python tools/run_layer.py --layers $L --iters 10 2>&1 | grep -v "^$" | sed -E "s/\{.*'tile_n': ([0-9]+).*'stages': ([0-9]+).*\}/bn=\1 st=\2/"; }
run FVY_DBG=1
