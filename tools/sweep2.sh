fmt() { grep -oE "^[0-9]+ |'idx': [0-9]+|'stages': [0-9]+|[0-9.]+ us" | paste -sd' ' | sed 's/ us /us\n/g'; }
L=11,69,28,27,45,9
echo "== default"; python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== NB=6"; FVY_NB=6 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== NB=5 LEAD=3 NB_RES=5"; FVY_NB=5 FVY_NB_RES=5 FVY_LEAD=3 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== GROUPS_KN=150"; FVY_GROUPS_KN=150 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== GROUPS_KN=150 NB=4"; FVY_GROUPS_KN=150 FVY_NB=4 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== GROUPS_KN=300 NB=4"; FVY_GROUPS_KN=300 FVY_NB=4 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
echo "== CTA2=0 NB=6"; FVY_CTA2=0 FVY_NB=6 python tools/run_layer.py --layers $L --iters 5 2>&1 | fmt
