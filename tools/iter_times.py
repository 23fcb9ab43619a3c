"""Per-iteration decode times (CUDA events per decode + host enqueue time) and SM clocks while it runs.
python tools/iter_times.py [B] [latent] [iters]"""
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from vae_decode_hdr_b200.engine import HdrVaeEngine  # noqa: E402
from vae_decode_hdr_b200.synthetic import random_decoder_state_dict, synthetic_latent  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dev = torch.device("cuda:0")
eng = HdrVaeEngine(random_decoder_state_dict(0), dev)
z = synthetic_latent(B, L, L).to(dev)
clk = []
stop = False


def sample():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        clk.append((time.perf_counter(), line.strip()))
        if stop:
            break
    p.terminate()


th = threading.Thread(target=sample, daemon=True)
th.start()
for _ in range(2):
    eng.decode(z, "moderate", want_stats=False)
torch.cuda.synchronize()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
host = []
t_begin = time.perf_counter()
for i in range(iters):
    t0 = time.perf_counter()
    evs[i][0].record()
    eng.decode(z, "moderate", want_stats=False)
    evs[i][1].record()
    host.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
t_end = time.perf_counter()
stop = True
gpu = [a.elapsed_time(b) for a, b in evs]
print("gpu ms per decode :", " ".join(f"{g:6.1f}" for g in gpu))
print("host enqueue ms   :", " ".join(f"{h:6.1f}" for h in host))
print(f"wall {1e3 * (t_end - t_begin) / iters:.1f} ms/decode over {iters} back-to-back decodes")
inside = [c for t, c in clk if t_begin <= t <= t_end]
print("clock samples during the loop (sm MHz, W, sw_power_cap, hw_slowdown, sw_thermal):")
for c in inside[:: max(1, len(inside) // 12)]:
    print("   ", c)
# same thing with a device sync after every decode (no queue build-up)
ts = []
for i in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.decode(z, "moderate", want_stats=False)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("synced wall ms    :", " ".join(f"{t:6.1f}" for t in ts))
