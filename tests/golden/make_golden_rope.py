"""Generate tests/golden/rope_golden.npz by running transformers' OWN rotary-embedding code.

Run in the build container only:

    python tests/golden/make_golden_rope.py

The reference applies RoPE through ``apply_rotary_pos_emb`` / ``LlamaRotaryEmbedding`` imported from
transformers==4.37.2 (hydragen/llama.py:1-10, 47-55, 494-501; requirements.txt) -- third-party code that is
not under /root/reference.  The installed transformers (5.5) keeps the same arithmetic
(``q * cos + rotate_half(q) * sin``, half-split rotation); since v4.38 the ``cos[position_ids]`` gather is
done by the caller, which is what this script does, exactly as 4.37.2's own function body did.  The cached
tables follow 4.37.2's ``_set_cos_sin_cache`` (inv_freq = base^(-2i/d), emb = cat(freqs, freqs)), cast to
the activation dtype as HydragenLlamaRotaryEmbedding.forward does.

Stored per case: the rotated q and k (bit patterns: 16-bit outputs are saved as int16 views) and a float64
checksum of the regenerated inputs.  tests/test_oracle.py checks oracle/hydragen_oracle.py against these on
CPU; tests/test_rope_gpu.py checks the CUDA kernel against them on the GPU -- both BIT-EXACT.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}

# name, b, s, hq, hkv, d, dtype, max_pos, seed
CASES = [
    ("decode_7b_slice", 16, 1, 32, 32, 128, "bfloat16", 4096, 0),
    ("decode_gqa_fp16", 9, 1, 8, 2, 128, "float16", 2048, 1),
    ("prefill_d64", 2, 37, 4, 4, 64, "bfloat16", 512, 2),
    ("fp32_d32", 3, 5, 6, 3, 32, "float32", 256, 3),
    ("mqa_d256", 5, 2, 4, 1, 256, "float16", 1024, 4),
]


def make_inputs(b, s, hq, hkv, d, dtype, max_pos, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(b, s, hq, d, generator=g).to(DT[dtype])
    k = torch.randn(b, s, hkv, d, generator=g).to(DT[dtype])
    pos = torch.randint(0, max_pos, (b, s), generator=g)
    return q, k, pos


def checksum(q, k, pos) -> float:
    return float(q.double().sum() + 3.0 * k.double().sum() + 1e-3 * pos.double().sum())


def tables_4_37(dim, max_pos, base, dtype):
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    t = torch.arange(max_pos, dtype=torch.int64).type_as(inv_freq)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def main():
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb

    out = {}
    for name, b, s, hq, hkv, d, dtype, max_pos, seed in CASES:
        q, k, pos = make_inputs(b, s, hq, hkv, d, dtype, max_pos, seed)
        cos, sin = tables_4_37(d, max_pos, 10000.0, DT[dtype])
        qe, ke = apply_rotary_pos_emb(q, k, cos[pos], sin[pos], unsqueeze_dim=2)
        view = (lambda t: t.view(torch.int16).numpy()) if DT[dtype] != torch.float32 else (lambda t: t.numpy())
        out[name + "/q"] = view(qe.contiguous())
        out[name + "/k"] = view(ke.contiguous())
        out[name + "/checksum"] = np.float64(checksum(q, k, pos))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rope_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
