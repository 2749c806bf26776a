L=${LAYERS:-1,2,3,6,10,11,27,28,45,58}
run() { echo "== $*"; env "$@" python tools/run_layer.py --layers $L --iters 5 2>&1 | grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*\}/st=\1/"; }
run FVY_DBG=1
run FVY_NB=4 FVY_NB_RES=6
run FVY_NB=2 FVY_NB_RES=3
