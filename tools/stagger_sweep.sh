#!/bin/bash
# sweep the start-up stagger of the ring forward kernel (debug knob)
for ns in 0 1500 3000 4500 7000; do
  echo "stagger $ns"; HMMCUDA_STAGGER_NS=$ns python tools/quickbench.py 18000000 2>&1 | sed -n "3,3p"
done
