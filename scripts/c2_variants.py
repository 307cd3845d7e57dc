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
HotPath.coarse_on_one_stream = True
run("both coarse scales on one side stream")
HotPath.coarse_on_one_stream = False
ops.KEEP_E = False
run("backward recomputes S / E instead of reading what the forward kept")
ops.KEEP_E = True
_lib.lib().dcnet_gemm_select(7)
run("4 epilogue warps in the fp16 exp / dS epilogues")
_lib.lib().dcnet_gemm_select(0)
ops.BWD_FP16 = False
run("co-attention backward on tf32 operands (round-2a)")
ops.BWD_FP16 = True
ops.RN_TF32 = False
run("operands truncated by the MMA instead of rounded by their producers (round 1)")
ops.RN_TF32 = True
