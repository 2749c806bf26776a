for t in 1024 512 256 128; do
  echo "== FVY_SWEEP_THREADS=$t"
  FVY_SWEEP_THREADS=$t python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('bench value %.0f post_ms %.4f' % (j['value'], j['roofline']['postprocess_ms']))"
  FVY_SWEEP_THREADS=$t python tools/post_bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('stress ms %.3f' % j['ms_per_step'])"
done
