import sys, os, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = bench.reference_function_leg(dev)
print(json.dumps({k: {a: b for a, b in v.items() if "diff" in a or "ms" in a} for k, v in out.items()}, indent=1))
