"""Multi-GPU check (torchrun, >= 2 GPUs): head-axis tensor-parallel generate() on the CUDA kernels + NVLS all-reduce
== the unsharded model on one GPU (teacher-forced logits), with and without CUDA-graph decode."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config  # noqa: E402
from hydragen_b200.tp import from_config_tp, shard_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
cfg = dict(hidden_size=1024, intermediate_size=2048, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=8, vocab_size=512, max_position_embeddings=1024)


def gen(model, graph, overrides=None):
    g = torch.Generator().manual_seed(0)
    ids = [torch.randint(3, 500, (1, 300), generator=g).to(dev), torch.randint(3, 500, (2, 9), generator=g).to(dev)]
    model.setup_caches(max_unique_batch_size=8, max_unique_seq_length=6, max_shared_batch_sizes=[1, 2], max_shared_seq_lengths=[300, 9])
    model.graph(graph)
    return model.generate(input_ids=ids, num_return_sequences=4, max_new_tokens=6, temperature=0.0, return_logits=True, token_overrides=overrides)


full = HydragenLlamaForCausalLM.from_config(llama_config("tiny", **cfg), dtype=torch.bfloat16, device=dev, seed=0, init_std=0.05)
ref_ids, ref_logits = gen(full, False)
tp = from_config_tp(llama_config("tiny", **cfg), dtype=torch.bfloat16, device=dev, seed=0)
tp.load_state_dict(shard_state_dict(full.state_dict(), rank, world), strict=False)
ok = True
for graph in (False, True):
    ids, logits = gen(tp, graph, overrides=ref_ids)
    a, b = torch.stack(logits).float(), torch.stack(ref_logits).float()
    err, scale = (a - b).abs().max().item(), b.abs().max().item()
    ok = ok and err < 0.05 * scale
    if rank == 0:
        from hydragen_b200.tp import _AllReduce

        nv = any(v is not None for v in _AllReduce._arenas.values())
        print(f"tp={world} graph={graph}: max |logit diff| {err:.3e} (max |logit| {scale:.2f}), NVLS all-reduce in use: {nv}", flush=True)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("tp parity", "ok" if ok else "FAILED", flush=True)
os._exit(0 if ok else 1)
