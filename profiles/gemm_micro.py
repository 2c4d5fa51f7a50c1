"""Graph-replayed micro-benchmark of the tcgen05 linear kernel: fixed overhead vs K (python profiles/gemm_micro.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vaura_b200 import _cabi  # noqa: E402

lib = _cabi.load()
torch.cuda.set_device(0)
side = torch.cuda.Stream()


def bench(R, N, K, bn, nmat=24, reps=20):
    W = (torch.randn(nmat, N, K, device="cuda") * 0.05).to(torch.bfloat16)
    A = torch.randn(R, K, device="cuda").to(torch.bfloat16)
    y = torch.empty(R, N, device="cuda")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for _ in range(2):
            for l in range(nmat):
                _cabi.check(lib.vaura_linear_bf16(A.data_ptr(), W[l].data_ptr(), y.data_ptr(), R, N, K, bn, side.cuda_stream), "lin")
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for l in range(nmat):
                _cabi.check(lib.vaura_linear_bf16(A.data_ptr(), W[l].data_ptr(), y.data_ptr(), R, N, K, bn, side.cuda_stream), "lin")
        g.replay()
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            g.replay()
        e1.record(side)
        side.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * nmat)
    mb = N * K * 2 / 1e6
    print(f"R={R:4d} N={N:5d} K={K:5d} bn={bn:3d}: {us:7.2f} us/launch  weights {mb:6.2f} MB -> {mb / us * 1e3:7.1f} GB/s", flush=True)


for K in (64, 256, 512, 1536):
    bench(64, 8192, K, 64)
for bn in (32, 64, 128):
    bench(64, 4608, 1536, bn)
bench(16, 8192, 1536, 64)
bench(128, 8192, 1536, 64)
bench(64, 1536, 4096, 64)
