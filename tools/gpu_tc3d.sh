#!/bin/bash
for cfg in "1 0" "2 0" "1 4" "2 4"; do set -- $cfg
echo "=== CL=$1 DBG=$2"; MDIL_TC_CLUSTER=$1 MDIL_TC3_DBG=$2 timeout -s KILL 60 python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider -k "full_size_shapes" 2>&1 | grep -E "passed|failed|Error|assert" | head -5
done
