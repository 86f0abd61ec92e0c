run() { python bench.py $1 --steps 3 --warmup 2 --no-cpu --no-e2e --flags $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$3', round(d['value']/1e9,3), round(d['roofline']['frac'],3))"; }
for MB in 4 5 6; do run "" $((32*256 + MB*65536)) "PL G32 MB$MB"; done
run "" $((16*256 + 5*65536)) "PL G16 MB5"
for MB in 4 5 6; do run "--workload er-100k-1M-sparseotf" $((16*256 + MB*65536)) "ER G16 MB$MB"; done
run "--workload er-100k-1M-sparseotf" $((32*256 + 5*65536)) "ER G32 MB5"
