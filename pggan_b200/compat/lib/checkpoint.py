from pggan_b200.checkpoint import *  # noqa: F401,F403
