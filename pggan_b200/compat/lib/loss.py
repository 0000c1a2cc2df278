from pggan_b200.loss import *  # noqa: F401,F403
