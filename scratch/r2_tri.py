import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
dev = torch.device("cuda:0")
r = bench.roofline_at_scale(dev, 6545.6)
print("NSVF_TRI_BWD=%s" % os.environ.get("NSVF_TRI_BWD", "1"))
for k, v in r.items():
    print("  ", k, v if isinstance(v, str) else (v["ms"], v["frac"]))
