from pggan_b200.blocks import *  # noqa: F401,F403
