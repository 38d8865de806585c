run() { PRIFIT_GRAPH_PRIO=$1 PRIFIT_ROWS_CLUSTER=$2 PRIFIT_GRAPH_BRANCHES=${3:-3} timeout 300 python bench.py --steps 30 --warmup 5 --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
for cfg in "0 4 3" "1 4 3" "0 4 2" "0 4 4" "0 8 3" "0 4 3"; do set -- $cfg
  echo "prio=$1 cluster=$2 branches=$3: $(run $1 $2 $3)"
done
