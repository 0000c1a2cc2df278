from pggan_b200.utils import *  # noqa: F401,F403
