"""kernel timeline of one CUDA-graph replay of the C2 step (torch.profiler / CUPTI): per-stream busy time, critical-path hints."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, random
from dcnet_b200 import synth
from dcnet_b200.hotpath import HotPath
dev = torch.device("cuda")
pairs, size = (int(sys.argv[2]) if len(sys.argv) > 2 else 8), (int(sys.argv[1]) if len(sys.argv) > 1 else 256)
tag = sys.argv[3] if len(sys.argv) > 3 else "timeline"
B = 2 * pairs
synth.seed_all(13)
hp = HotPath(size).to(dev).train()
g = torch.Generator().manual_seed(9000)
b = synth.make_hotpath_batch(pairs, size, g)
grad_keys = ('raw', 'flang', 'fa', 'context', 'head', 'loc')
flat = dict(raw=b['raw'], flang=[b['flang']], fa=[b['fa']], context=[b['context']], head=b['head'], loc=b['loc'], dy_head=b['dy_head'], bbox=[b['bbox']])
static = {k: [t.to(dev).requires_grad_(k in grad_keys) for t in v] for k, v in flat.items()}
np_, ni_ = hp.draw_indices(B)
s_negpos, s_negidx = torch.from_numpy(np_).to(dev), torch.from_numpy(ni_).to(dev)
def run_step():
    return hp.step(static['raw'], static['flang'][0], static['fa'][0], static['context'][0], static['head'], static['loc'], static['dy_head'], static['bbox'][0], s_negpos, s_negidx)
def clear():
    for p in hp.parameters(): p.grad = None
    for k in grad_keys:
        for t in static[k]: t.grad = None
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        clear(); run_step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
clear()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    res = run_step()
for _ in range(5): graph.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    graph.replay(); torch.cuda.synchronize()
    graph.replay(); torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/timeline.json")
ev = json.load(open("/tmp/timeline.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
# second replay only
half = len(ks) // 2
ks = ks[half:]
t0 = ks[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ks)
print("kernels in one replay: %d, span %.1f us" % (len(ks), t1 - t0))
per = collections.defaultdict(float)
for e in ks: per[e["args"].get("stream")] += e["dur"]
for s, d in per.items(): print("  stream %s busy %.1f us" % (s, d))
with open("gpurun_out/%s.txt" % tag, "w") as f:
    for e in ks:
        f.write("%9.1f %8.1f s%-3s %s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:90]))
