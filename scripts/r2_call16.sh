#!/bin/bash
for m in 6 2 4 7 0; do RT_SORT_SEGS=$m python scripts/r2_probe.py own 2>/dev/null | head -1 | sed "s/^/segs=$m /"; done
