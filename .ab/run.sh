#!/bin/bash
cp tensortrains.jl_b200/ttb200/libttb200.so .ab/lib_cur.so
for v in "$@"; do
  cp .ab/lib_$v.so tensortrains.jl_b200/ttb200/libttb200.so
  for i in 1 2; do
    timeout 200 python bench.py --no-extra --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['roofline']['kernels']; print('$v', round(d['ms_per_step'],2), {n: round(k[n]['ms'],2) for n in list(k)[:4]})"
  done
done
cp .ab/lib_cur.so tensortrains.jl_b200/ttb200/libttb200.so
