SR4D_TC_DEBUG=1 timeout -s KILL 300 python tools/train_once.py 8 1 2>&1 | grep "wgrad2 dbg" | grep "D=24" | head -3
