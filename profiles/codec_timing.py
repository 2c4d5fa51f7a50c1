"""Times the DAC decode (tcgen05 implicit-GEMM path vs SIMT path) on one B200 and checks both against the
CPU oracle.  Usage: python profiles/codec_timing.py [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.dac_oracle import DacDecodeOracle  # noqa: E402
from vaura_b200.codec import DacModelWrapper  # noqa: E402
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict  # noqa: E402
from vaura_b200.weights import codec_flops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = make_codec_state_dict(FULL_CODEC, 100)
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, 1024, (B, 9, 220), generator=g)
ref = DacDecodeOracle(sd, FULL_CODEC).decode(codes[:1])
for mode in ("tc", "simt"):
    os.environ["VAURA_CODEC_SIMT"] = "1" if mode == "simt" else "0"
    m = DacModelWrapper(44100, dims=FULL_CODEC)
    m.load_state_dict(sd, device="cuda:0")
    wav = m.decode(codes.cuda())
    torch.cuda.synchronize()
    err = (wav[:1].float().cpu() - ref)
    snr = 10 * torch.log10(ref.pow(2).sum() / err.pow(2).sum())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.decode(codes.cuda())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{mode}: B={B} {ms:.2f} ms/batch  {codec_flops(FULL_CODEC, 220) * B / ms / 1e9:.1f} TFLOP/s  SNR vs fp32 oracle {snr:.1f} dB",
          flush=True)
