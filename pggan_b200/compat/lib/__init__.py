"""Import shims: put `pggan_b200/compat` on sys.path and the reference's own module paths
(`lib.layers`, `lib.custom_layers`, `lib.blocks`, `lib.utils`, `lib.loss`, `lib.checkpoint`, `pggan.nets`,
`pggan.loss`, `pggan.model`) resolve to the B200-native implementations, so reference scripts that do
`from lib.layers import EqualizedConv2d` / `from pggan.nets import Generator` run unchanged."""
