#!/bin/bash
SR4D_TC_DEBUG=1 timeout 200 python tools/fwd_once.py 8 2 2>&1 | grep "tc dbg" | tail -30 | sort | uniq -c | sort -rn | awk '{$1="";print}' | sort -u -k3,6 | head -6
