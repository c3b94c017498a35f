python __graft_entry__.py --smoke 2>&1 | tail -8
MDIL_PAIR_IMPL=ffma python __graft_entry__.py --smoke 2>&1 | tail -3
MDIL_WGRAD_IMPL=ffma python __graft_entry__.py --smoke 2>&1 | tail -3
