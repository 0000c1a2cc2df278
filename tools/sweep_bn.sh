for bn in 32 64 128 256; do echo "== BN $bn"; PGGAN_FORCE_BN=$bn timeout 100 python tools/gemm_micro.py 10 2>&1 | grep conv_fprop; done
