"""DAC encode on the GPU against the CPU oracle (oracle/dac_oracle.py: DacEncodeOracle, fp32).  The GPU path
stores activations in fp16 (the precision the reference itself runs the codec in, vaura_model.py:92) and quantises with fp32
residuals; codes are discrete, so parity is stated as: latent SNR >= 50 dB, first-codebook agreement >= 97 %, and every cell
whose earlier codebooks agree either has the oracle's code or sits within 0.02 of a tie in the oracle's own similarity."""
import pytest
import torch

from oracle.dac_oracle import DacEncodeOracle
from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict

pytestmark = pytest.mark.gpu


def snr_db(ref, x):
    ref, x = ref.double().flatten(), x.double().flatten()
    return float(10 * torch.log10(ref.pow(2).sum() / (ref - x).pow(2).sum().clamp_min(1e-30)))


@pytest.fixture(scope="module")
def codec():
    from vaura_b200.codec import DacModelWrapper

    m = DacModelWrapper(model_sr=44100, dims=FULL_CODEC)
    m.load_state_dict(make_codec_state_dict(FULL_CODEC, 100, with_encoder=True), device="cuda:0")
    return m


def test_encode_matches_oracle(codec):
    sd = make_codec_state_dict(FULL_CODEC, 100, with_encoder=True)
    oracle = DacEncodeOracle(sd, FULL_CODEC)
    g = torch.Generator().manual_seed(9)
    L = 44100 + 37  # not a multiple of the hop: exercises DAC.preprocess
    t = torch.arange(L) / 44100.0
    wav = torch.stack([0.3 * torch.sin(2 * torch.pi * f0 * t) + 0.1 * torch.randn(L, generator=g) for f0 in (220.0, 1330.0, 57.0)])[:, None]
    codes, latent = codec.encode(wav.cuda(), _return_latent=True)
    z = oracle.encode_latent(oracle.preprocess(wav))
    ref, margin = oracle.quantize(z, return_margins=True)
    assert codes.shape == ref.shape == (3, 9, 87) and codes.dtype == torch.int64
    s = snr_db(z.transpose(1, 2), latent.float().cpu())
    agree = (codes.cpu() == ref)
    print(f"[dac encode] latent SNR {s:.1f} dB; agreement per codebook {[round(float(a), 3) for a in agree.float().mean(dim=(0, 2))]}")
    assert s > 50.0, s
    assert float(agree[:, 0].float().mean()) >= 0.97
    prefix_ok = torch.cumprod(torch.cat([torch.ones_like(agree[:, :1]), agree[:, :-1]], dim=1).int(), dim=1).bool()
    bad = prefix_ok & ~agree & (margin > 0.02)
    assert not bad.any(), f"{int(bad.sum())} cells differ from the oracle away from a tie"


def test_encode_decode_round_trip_and_shapes(codec):
    """encode -> decode keeps the frame count (models/modules/dac/model.py:30-48); (L,) and (1, L) inputs are accepted."""
    wav = 0.2 * torch.randn(512 * 20, generator=torch.Generator().manual_seed(1))
    c1 = codec.encode(wav.cuda())
    c2 = codec.encode(wav[None].cuda())
    c3 = codec(wav[None, None].cuda())
    assert c1.shape == (1, 9, 20) and torch.equal(c1, c2) and torch.equal(c1, c3)
    assert int(c1.min()) >= 0 and int(c1.max()) < 1024
    audio = codec.decode([(c1, None)])
    assert audio.shape == (1, 1, 512 * 20)
    with pytest.raises(ValueError):
        codec.encode(torch.zeros(1, 2, 1024).cuda())


def test_encode_of_less_than_one_hop_and_batches_beyond_the_chunk(codec):
    """Edge sizes: 100 samples (padded to one 512-sample frame) and a batch larger than the 8-clip workspace chunk."""
    sd = make_codec_state_dict(FULL_CODEC, 100, with_encoder=True)
    oracle = DacEncodeOracle(sd, FULL_CODEC)
    g = torch.Generator().manual_seed(4)
    short = 0.3 * torch.randn(1, 1, 100, generator=g)
    c = codec.encode(short.cuda())
    assert c.shape == (1, 9, 1)
    assert int(c[0, 0, 0]) == int(oracle.encode(short)[0, 0, 0])
    wav = 0.3 * torch.randn(11, 1, 512 * 6, generator=g)
    many = codec.encode(wav.cuda())
    one = torch.cat([codec.encode(wav[i:i + 1].cuda()) for i in range(11)])
    assert many.shape == (11, 9, 6) and torch.equal(many, one)
