import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EULER_KNOBS = dict(num_steps=8, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0,
                   truncation_factor=0.8, rescale_k=1.2, rescale_sigma=3.0, speaker_kv_scale=1.5,
                   speaker_kv_max_layers=2, speaker_kv_min_t=0.6)
PLAIN_KNOBS = dict(EULER_KNOBS, truncation_factor=None, rescale_k=None, rescale_sigma=None, speaker_kv_scale=None,
                   speaker_kv_max_layers=None, speaker_kv_min_t=None)
HANDLER_KNOBS = dict(num_steps=40, cfg_scale_text=3.0, cfg_scale_speaker=8.0, cfg_min_t=0.5, cfg_max_t=1.0,
                     truncation_factor=None, rescale_k=None, rescale_sigma=None, speaker_kv_scale=None,
                     speaker_kv_max_layers=None, speaker_kv_min_t=None)


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def gold(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=True)


def byte_tokens(prompts, max_length):
    """Byte tokenizer + padding exactly as reference inference.py:115-136,192-214 (BOS 0 + UTF-8 bytes, prefix mask)."""
    ids = torch.zeros(len(prompts), max_length, dtype=torch.int32)
    mask = torch.zeros(len(prompts), max_length, dtype=torch.bool)
    for i, p in enumerate(prompts):
        b = [0] + list(p.encode("utf-8"))
        n = min(len(b), max_length)
        ids[i, :n] = torch.tensor(b[:n], dtype=torch.int32)
        mask[i, :n] = True
    return ids, mask
