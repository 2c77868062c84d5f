for c in "" "192,384,460" "160,320,440" "200,400,480" "240,440" "128,256,384,470" "250,450,490"; do
  echo "cuts=[$c]"; SKB_SUB_CUTS="$c" python bench.py --skip-cpu-baseline --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
done
