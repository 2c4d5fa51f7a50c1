"""One prefill window of a chunked long clip (B = 1, prompt of 166 tokens) for an ncu launch list: generate() stopped
after the first sampled column."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vaura_b200.synthetic import FULL_CODEC, FULL_SAMPLER, build_model, make_avclip_features  # noqa: E402

m = build_model(FULL_SAMPLER, FULL_CODEC)
feats = make_avclip_features(1, 3).cuda()
prompt = torch.randint(0, 1024, (1, 9, 166)).cuda()
m.generate(frames=feats, audio=prompt, max_new_tokens=221, prompt_is_encoded=True, _decode_audio=False, _end_offset=168,
           use_sampling=True, top_k=128)
torch.cuda.synchronize()
print("done")
