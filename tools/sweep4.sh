echo "== nowork pdl"; FVY_NOWORK=1 python tools/run_layer.py --layers 28,27,10,2 --iters 20 2>&1 | grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*'tiles': ([0-9]+).*\}/st=\1 tiles=\2/"
echo "== nowork no pdl"; FVY_PDL=0 FVY_NOWORK=1 python tools/run_layer.py --layers 28,27,10,2 --iters 20 2>&1 | grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*'tiles': ([0-9]+).*\}/st=\1 tiles=\2/"
echo "== work no pdl"; FVY_PDL=0 python tools/run_layer.py --layers 28,27,10,2 --iters 20 2>&1 | grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*'tiles': ([0-9]+).*\}/st=\1 tiles=\2/"
echo "== work pdl"; python tools/run_layer.py --layers 28,27,10,2 --iters 20 2>&1 | grep -v "^$" | sed -E "s/\{.*'stages': ([0-9]+).*'tiles': ([0-9]+).*\}/st=\1 tiles=\2/"
