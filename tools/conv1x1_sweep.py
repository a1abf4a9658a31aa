"""CUDA-event timings of the teacher's HBM-bound 1x1 convolutions under the kernel's process-global switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_cases as nc

L = nc.L
names = ["conv1x1_256_1024_res", "conv1x1_1024_256", "conv1x1_512_2048_res"]
bytes_ = {"conv1x1_256_1024_res": 8 * 65 * 129 * 4 * (256 + 2 * 1024), "conv1x1_1024_256": 8 * 65 * 129 * 4 * (1024 + 256),
          "conv1x1_512_2048_res": 8 * 65 * 129 * 4 * (512 + 2 * 2048)}
configs = [("default (in-place res)", lambda: None), ("staged residual ring", lambda: L.skd_set_conv_res_prefetch(1)), ("pairs forced", lambda: L.skd_set_conv_cta_pairs(3)), ("pairs off", lambda: L.skd_set_conv_cta_pairs(0)),
           ("no residual ring", lambda: L.skd_set_conv_res_prefetch(0)), ("back-to-front tiles", lambda: L.skd_set_conv_tile_order(1))]
for label, setup in configs:
    L.skd_set_conv_cta_pairs(1); L.skd_set_conv_res_prefetch(2); L.skd_set_conv_tile_order(0)
    setup()
    for n in names:
        us = nc.time_case(n)
        print("%-24s %-24s %7.1f us  %5.2f TB/s" % (label, n, us, bytes_[n] / us / 1e6), flush=True)
