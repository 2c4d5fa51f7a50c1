"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py [avclip|encode|decode|prefill|prefill_bf16|step]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
what = sys.argv[1] if len(sys.argv) > 1 else "avclip"
if what == "avclip":
    from vaura_b200.features import MotionFormer
    from vaura_b200.synthetic import make_motionformer_state_dict, make_video_segments
    m = MotionFormer(extract_features=True)
    m.load_state_dict(make_motionformer_state_dict(7), device="cuda:0")
    out, _ = m(make_video_segments(1, 3, 2).cuda())
    print("avclip", tuple(out.shape), float(out.abs().max()))
elif what in ("encode", "decode"):
    from vaura_b200.codec import DacModelWrapper
    from vaura_b200.synthetic import FULL_CODEC, make_codec_state_dict
    c = DacModelWrapper(44100, dims=FULL_CODEC)
    c.load_state_dict(make_codec_state_dict(FULL_CODEC, 100, with_encoder=True), device="cuda:0")
    if what == "encode":
        codes = c.encode((0.3 * torch.randn(2, 1, 512 * 9 + 17)).cuda())
        print("encode", tuple(codes.shape), int(codes.max()))
    else:
        wav = c.decode(torch.randint(0, 1024, (2, 9, 7)).cuda())
        print("decode", tuple(wav.shape), float(wav.float().abs().max()))
elif what == "step":  # decode_step_fused_bf16 (16 rows, a few columns) and decode_step_cluster (1 row)
    from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
    m = build_model(FULL_SAMPLER, FULL_CODEC)
    for B in (16, 1):
        o = m.generate(frames=make_avclip_features(B, 3).cuda(), max_new_tokens=3, prompt_is_encoded=True, _decode_audio=False,
                       use_sampling=True, top_k=64, return_sampled_indices=True)
        print("step", B, tuple(o["sampled_indices"].shape))
elif what == "prefill_bf16":  # prompt prefill of a sampling call: bf16 GEMM operands (transformer_pass_tc1)
    from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
    m = build_model(FULL_SAMPLER, FULL_CODEC)
    prompt = torch.randint(0, 1024, (1, 9, 150)).cuda()
    o = m.generate(frames=make_avclip_features(1, 3).cuda(), audio=prompt, max_new_tokens=160, prompt_is_encoded=True,
                   _decode_audio=False, _end_offset=153, use_sampling=True, top_k=64, return_sampled_indices=True)
    print("prefill_bf16", tuple(o["sampled_indices"].shape))
else:
    from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features
    m = build_model(FULL_SAMPLER, FULL_CODEC)
    prompt = torch.randint(0, 1024, (1, 9, 150)).cuda()
    o = m.generate(frames=make_avclip_features(1, 3).cuda(), audio=prompt, max_new_tokens=160, prompt_is_encoded=True,
                   _decode_audio=False, _end_offset=153, use_sampling=False, return_sampled_indices=True)
    print("prefill", tuple(o["sampled_indices"].shape))
torch.cuda.synchronize()
