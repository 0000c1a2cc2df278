from pggan_b200.trainer import ProgressiveGAN, make_args  # noqa: F401
