#!/usr/bin/env python
"""e2e leg of bench.py alone, over the number of host lanes (threads x handles)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", default="2,3,4,6")
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
torch, dist, world, rank, local, dev = bench.dist_setup()
import avoid_mpc_b200 as A
ns = argparse.Namespace(batch=a.batch, warm="ref", tol=1e-8, max_iter=50, steps=24)
for L in [int(x) for x in a.lanes.split(",")]:
    e = bench.e2e_leg(A, torch, dist, dev, local, world, ns, lanes_n=L)
    print(json.dumps({"lanes": L, "batch": a.batch, "f32": e["value"], "f32_GBps": e["h2d_GBps"], "u16": e["depth_u16"]["value"],
                      "u16_GBps": e["depth_u16"]["h2d_GBps"], "cloud_upload": e["cloud_upload"]["value"],
                      "alone_GBps": e["pinned_h2d_GBps_alone"]}), flush=True)
