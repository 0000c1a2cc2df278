from pggan_b200.nets import *  # noqa: F401,F403
from pggan_b200.nets import Generator, Discriminator  # noqa: F401
