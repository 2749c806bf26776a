f() { grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*'tiles': ([0-9]+).*\}/st=\1 tiles=\2/"; }
echo "== nowork=2 (bare launch)"; FVY_NOWORK=2 python tools/run_layer.py --layers 28,10,2 --iters 50 2>&1 | f
echo "== nowork=2 no pdl"; FVY_PDL=0 FVY_NOWORK=2 python tools/run_layer.py --layers 28,10,2 --iters 50 2>&1 | f
echo "== nowork=1"; FVY_NOWORK=1 python tools/run_layer.py --layers 28,10,2 --iters 50 2>&1 | f
