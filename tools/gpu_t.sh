MDIL_TC_TRACE=1 timeout 60 python tools/trace_tc.py 2>&1 | grep wgrad_tc | sed -n '1p;2p;7p;8p' | cut -c1-330
bash tools/gpu_quick.sh ncu
