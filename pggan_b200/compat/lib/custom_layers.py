from pggan_b200.custom_layers import *  # noqa: F401,F403
