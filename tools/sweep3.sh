L=${LAYERS:-28,11}
run() { echo "== $*"; env "$@" python tools/run_layer.py --layers $L --iters 10 2>&1 | grep -v "^$" | sed -E "s/\{.*'tile_n': ([0-9]+).*'stages': ([0-9]+).*\}/bn=\1 st=\2/" | grep -v "epilogue group"; }
run FVY_DBG=1 FVY_TAIL_SPLIT=1
run FVY_DBG=1 FVY_TAIL_SPLIT=0
