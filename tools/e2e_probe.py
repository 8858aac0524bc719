"""Where does the end-to-end path lose time?  Pipelined steps with device-resident results, + host copies, + CSV text."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
from octa_autosegmentation_b200.config import default_config
from octa_autosegmentation_b200.pipeline import Pipeline

pipe = Pipeline(default_config(), volume_dims=(1216, 1216, 16))
K, SB, L = int(os.environ.get("K", 12)), 32, 8
seed = [1_000_000]
def batches(k):
    out = []
    for _ in range(2 * k):
        out.append(list(range(seed[0], seed[0] + SB))); seed[0] += SB
    return out
def run(k, d2h, csv):
    for _ in pipe.run_pipelined(batches(k), d2h=d2h, csv=csv, in_flight=L):
        pass
    torch.cuda.synchronize()
run(7, False, False); run(7, True, True)
for name, d2h, csv in (("device resident", False, False), ("+ D2H label/image", True, False), ("+ CSV text", True, True), ("device resident", False, False)):
    torch.cuda.synchronize(); t = time.time(); run(K, d2h, csv); dt = time.time() - t
    print("%-22s %.1f graphs/s (%.1f ms per 64)" % (name, K * 64 / dt, dt / K * 1e3))
