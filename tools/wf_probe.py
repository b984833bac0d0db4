import sys, os, json, torch
sys.path.insert(0, '/root/repo')
import bench
hbm, _, _ = bench.load_peaks()
torch.cuda.set_device(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
print(json.dumps(bench.bench_waterfall(torch, torch.device('cuda', 0), hbm)))
