for d in 0 1 2 3; do echo "== DBG=$d"; MDIL_TC3_DBG=$d MDIL_TC_TRACE=1 timeout -s KILL 120 python tools/trace_tc.py 2>&1 | grep -E "pair_tc3" | sed -n '1p;3p;5p;7p' | cut -c1-330; done
