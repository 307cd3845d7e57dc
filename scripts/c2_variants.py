"""step time of C2 / C3 (CUDA-graph replay, L2 flushed) under switches of the step's composition"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dcnet_b200 import ops, _lib
from dcnet_b200.hotpath import HotPath
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
key = sys.argv[1] if len(sys.argv) > 1 else "c2"


def run(name):
    m = bench.measure_hotpath(key, 20, 3, 0, 1, 0, dev, None)
    print("%-50s %.4f ms/step  e2e %.4f  launches %d" % (name, m["ms_per_step"], m["e2e"]["ms_per_step"], m["launches_per_step"]), flush=True)


run("default")
HotPath.side_priority = (-1, -1)
HotPath.aux_priority = (-2, -2)
run("coarse scales at priority -1, aux -2, finest 0")
HotPath.side_priority = (-1, -1)
HotPath.aux_priority = (-1, -1)
run("coarse scales and aux at priority -1, finest 0")
HotPath.side_priority = (-2, -1)
HotPath.aux_priority = (-3, -3)
run("coarsest -2, middle -1, aux -3, finest 0")
HotPath.side_priority = (0, 0)
HotPath.aux_priority = (-1, -1)
