from pggan_b200.loss import WGANGPLoss, Loss, LossInterface  # noqa: F401
