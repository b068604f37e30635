#!/bin/bash
RT_TIMING=1 python scripts/e2e_probe.py 2>&1 | tail -12
RT_BVH_TIMING=1 python scripts/e2e_probe.py 2>&1 | grep bvh_build | tail -2
