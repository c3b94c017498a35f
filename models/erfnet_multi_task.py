"""Drop-in location of the reference's multi-task joint model: ``models/erfnet_multi_task.py`` (train_multi_task.py
loads ``args.model`` from ``models/``)."""
from mdil_ss_b200.erfnet_multi_task import *  # noqa: F401,F403
from mdil_ss_b200.erfnet_multi_task import (DownsamplerBlock, non_bottleneck_1d, Encoder, UpsamplerBlock, Decoder,  # noqa: F401
                                            Net)
