"""Read-only / write-only / copy HBM bandwidth with plain torch ops (roofline context for the layer kernels)."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 30  # 4 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
x.fill_(1.0); y.fill_(2.0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
t = timed(lambda: x.fill_(3.0)); out["write_only_GBs"] = 4 * n / t / 1e9
t = timed(lambda: x.sum()); out["read_only_GBs"] = 4 * n / t / 1e9
t = timed(lambda: y.copy_(x)); out["copy_GBs_read_plus_write"] = 8 * n / t / 1e9
t = timed(lambda: torch.add(x, y, out=y)); out["add_2r1w_GBs"] = 12 * n / t / 1e9
print(json.dumps(out))
