"""Exists so the drivers' ``assert os.path.exists(args.model + ".py")`` (train_new_task_step2.py:463) holds when
they are started from the repository root with ``--model erfnet_RA_parallel``."""
from mdil_ss_b200.erfnet_RA_parallel import *  # noqa: F401,F403
from mdil_ss_b200.erfnet_RA_parallel import Net  # noqa: F401
