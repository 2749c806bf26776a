L=${LAYERS:-1,3,4,6,2,5}
run() { echo "== $*"; env "$@" python tools/run_layer.py --layers $L --iters 10 2>&1 | grep -v "^$" | sed -E "s/\{.*'tile_n': ([0-9]+).*'stages': ([0-9]+).*\}/bn=\1 st=\2/" | grep -v "timeline"; }
run FVY_DBG=1
