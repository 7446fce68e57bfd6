"""Diagnostic: what moves data between two GPUs of the box, and how fast.  One process, devices 0 and 1: peer access,
copy-engine bandwidth (cudaMemcpyPeerAsync through torch), both directions at once; then the NCCL send/recv bandwidth
of two ranks is measured by tools/nccl_sendrecv_probe.py under torchrun."""
import json
import subprocess
import torch

res = {"n_gpus": torch.cuda.device_count()}
if res["n_gpus"] >= 2:
    res["peer_access_0_1"] = torch.cuda.can_device_access_peer(0, 1)
    n = 256 << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    a2 = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    b2 = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    s0, s1 = torch.cuda.Stream(device=0), torch.cuda.Stream(device=1)
    for name, both in (("one_direction", False), ("both_directions", True)):
        best = 1e9
        for rep in range(4):
            torch.cuda.synchronize(0); torch.cuda.synchronize(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s0):
                e0.record(s0)
                for _ in range(4):
                    b.copy_(a, non_blocking=True)
                e1.record(s0)
            if both:
                with torch.cuda.stream(s1):
                    for _ in range(4):
                        a2.copy_(b2, non_blocking=True)
            torch.cuda.synchronize(0); torch.cuda.synchronize(1)
            best = min(best, e0.elapsed_time(e1))
        res[f"copy_engine_gbs_{name}"] = round(4 * n / (best * 1e-3) / 1e9, 1)
try:
    res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout.splitlines()[:12]
except Exception as e:
    res["topo"] = repr(e)
print(json.dumps(res, indent=1))
