"""Whole-model decode throughput at BASELINE.json configs[2]: random-init Llama-2-7B (no checkpoints offline),
one shared prompt of 2048 tokens, 1024 completions, bf16, CUDA-graph decode -- measured the way the reference's
scripts/synth.py does (:36-79, 217-226): time generate(max_new_tokens = N) and generate(max_new_tokens = 1) with
CUDA events and divide the decoded tokens by the difference.  The projections / MLP / lm_head are stock cuBLAS
(out of scope of the hot path); `attention_share` comes from the same run with disable_attention=True
(hydragen/llama.py:433-437).  Prints one JSON object."""

import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed_generate(model, ids, nrs, n_new, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.generate(input_ids=ids, num_return_sequences=nrs, max_new_tokens=n_new, temperature=100.0, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run(model_name="llama-2-7b", batch=1024, prefix=2048, new_tokens=128, iters=2, layers=None):
    from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config

    over = {} if layers is None else {"num_hidden_layers": layers}
    cfg = llama_config(model_name, **over)
    t0 = time.time()
    model = HydragenLlamaForCausalLM.from_config(cfg, dtype=torch.bfloat16, device="cuda", seed=0)
    model.setup_caches(max_unique_batch_size=batch, max_unique_seq_length=new_tokens, max_shared_batch_sizes=[1], max_shared_seq_lengths=[prefix])
    model.graph(True)
    ids = torch.randint(3, 31000, (1, prefix), device="cuda")
    res = {"model": model_name, "layers": cfg.num_hidden_layers, "batch": batch, "prefix": prefix, "new_tokens": new_tokens, "setup_s": None}
    timed_generate(model, ids, batch, 4)  # warm-up: graph capture, cuBLAS handles
    res["setup_s"] = round(time.time() - t0, 1)
    full = min(timed_generate(model, ids, batch, new_tokens) for _ in range(iters))
    pre = min(timed_generate(model, ids, batch, 1) for _ in range(iters))
    noattn = min(timed_generate(model, ids, batch, new_tokens, disable_attention=True) for _ in range(iters))
    noattn_pre = min(timed_generate(model, ids, batch, 1, disable_attention=True) for _ in range(iters))
    steps = new_tokens - 1
    dec_ms, na_ms = full - pre, noattn - noattn_pre
    res.update(decode_tokens_per_s=batch * steps / (dec_ms / 1e3), ms_per_decode_step=dec_ms / steps, prefill_ms=pre,
               no_attention_tokens_per_s=batch * steps / (na_ms / 1e3), attention_share=max(0.0, 1 - na_ms / dec_ms))
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-2-7b")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--prefix", type=int, default=2048)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--layers", type=int, default=None)
    a = ap.parse_args()
    print(json.dumps(run(a.model, a.batch, a.prefix, a.new_tokens, layers=a.layers)))
