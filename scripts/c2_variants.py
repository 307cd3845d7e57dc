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
HotPath.finest_priority = -1
HotPath.aux_priority = (-2, -2)
run("finest chain on a priority -1 stream, aux -2, coarse scales 0")
HotPath.finest_priority = -2
HotPath.aux_priority = (-3, -3)
HotPath.side_priority = (-1, 0)
run("finest -2, aux -3, middle scale -1, coarsest 0")
HotPath.finest_priority = None
HotPath.aux_priority = (-1, -1)
HotPath.side_priority = (0, 0)
ops.BWD_FP16 = False
run("co-attention backward on tf32 operands (round-2a)")
ops.BWD_FP16 = True
