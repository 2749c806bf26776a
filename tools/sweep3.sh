L=${LAYERS:-28,27,10,3,1,6}
run() { echo "== $*"; env "$@" python tools/run_layer.py --layers $L --iters 10 2>&1 | grep -v "^$" | sed -E "s/\{.*'tile_n': ([0-9]+).*'stages': ([0-9]+).*\}/bn=\1 st=\2/" | grep -v "timeline\|dbg conv_[0-9]*:"; }
run FVY_DBG=1
