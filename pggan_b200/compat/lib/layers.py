from pggan_b200.layers import *  # noqa: F401,F403
