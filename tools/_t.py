import sys, json, torch
sys.path.insert(0, "/root/repo")
import bench, nndepth_b200 as nb
nb.load_library()
peak, _ = bench.measured_peak()
r = bench.time_fused_lookup(torch.device("cuda"), 64, 2, peak)
print(json.dumps(r["plain_lookup_corr1d_lookup_lean_kernel"]))
print(json.dumps({k: r[k] for k in ("us_per_launch_l2_flushed", "frac")}), json.dumps(r["smooth_field_skewed_layout"]))
