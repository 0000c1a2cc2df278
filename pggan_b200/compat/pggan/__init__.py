"""see pggan_b200/compat/lib/__init__.py"""
