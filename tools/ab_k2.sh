#!/bin/bash
# A/B of libcramore_cuda.so variants on the GPU box: tools/ab_k2.sh NAME... (cramore_b200/variants/libccu_NAME.so;
# "default" = the shipped library).  Per variant: the tiled-K2 parity tests, then K2 times at C2 / skewed C2 / C5.
out=gpurun_out/ab_k2.log
: > $out
for n in "$@"; do
  if [ "$n" = default ]; then unset CCU_LIB; else export CCU_LIB=$PWD/cramore_b200/variants/libccu_$n.so; fi
  echo "=== $n" >> $out
  if [ -z "$AB_SKIP_TESTS" ]; then
    timeout 300 python -m pytest tests/test_demux_gpu.py -m gpu -x -q 2>&1 | tail -2 >> $out
  fi
  timeout 90 python tools/prof_demux.py 20000 C2 2>&1 | tail -1 >> $out
  timeout 90 python tools/prof_demux.py 5000 C2 1.5 800 2>&1 | tail -1 >> $out
  timeout 90 python tools/prof_demux.py 2000 C5 2>&1 | tail -1 >> $out
done
grep -E "===|ms_score|passed|failed|rror" $out | sed -E "s/.*'ms_score': ([0-9.]+).*/  ms_score \1/"
