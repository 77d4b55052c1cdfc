"""Micro-benchmark of vadx_stft_power_tc_i16 on the FireRed shape (perf experiments)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vadx
from vadx import lib, tables, synth
l = lib.load()
S, L, T = 8192, 16000, 98
dev = torch.device("cuda")
x = torch.from_numpy(synth.synth_chunks_fast(S, L)).to(dev)
basis, first, nb = tables.interleaved_basis(400, 400, "povey", "v2")
img = torch.from_numpy(lib.pack_stft_basis_tc(basis, nb, 0.97, 1.0)).to(dev)
out = torch.empty((S * T, 204), device=dev)
def run():
    lib.check(l.vadx_stft_power_tc_i16(x.data_ptr(), L, L, S, T, 160, 400, img.data_ptr(), nb, out.data_ptr(), 204, lib.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print(f"stft_tc: {e0.elapsed_time(e1) / 5:.3f} ms  debug={os.environ.get('VADX_TC_DEBUG', '0')}")
