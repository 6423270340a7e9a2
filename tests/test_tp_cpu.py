"""CPU tests (gloo, world_size 2) of the multi-GPU path: head-axis tensor parallelism of
hydragen_b200/tp.py (the split of hydragen/tp.py:30-124 of the reference, which the reference never tests).
The attention calls are routed to the CPU oracle (tests/oracle_patch.py) -- on the GPU box the same code runs
the CUDA kernels over NCCL; here the subject is the host logic: weight sharding, local head counts, per-rank
caches, the all-reduce after o_proj / down_proj, and identical sampling on every rank."""

import os
import socket
import sys
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _patch_oracle():
    import hydragen_b200.llama as L
    import oracle_patch as P

    for name in ("flash_attention", "flash_attention_seqlen", "hydragen_attention", "hydragen_attention_decode", "kv_append", "apply_rotary_pos_emb"):
        setattr(L, name, getattr(P, name))


def _run_generate(model, nrs=3, max_new=5):
    g = torch.Generator().manual_seed(0)
    ids = [torch.randint(3, 500, (1, 24), generator=g), torch.randint(3, 500, (2, 7), generator=g)]
    model.setup_caches(max_unique_batch_size=2 * nrs, max_unique_seq_length=max_new, max_shared_batch_sizes=[1, 2], max_shared_seq_lengths=[24, 7])
    return model.generate(input_ids=ids, num_return_sequences=nrs, max_new_tokens=max_new, temperature=0.0, return_logits=True)


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        _patch_oracle()
        from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config
        from hydragen_b200.tp import from_config_tp

        cfg = llama_config("tiny")
        full = HydragenLlamaForCausalLM.from_config(llama_config("tiny"), dtype=torch.float32, device="cpu", seed=0, init_std=0.08)
        ref_ids, ref_logits = _run_generate(full)
        tp_model = from_config_tp(cfg, dtype=torch.float32, device="cpu", seed=0)
        # from_config_tp uses its own init_std: re-shard the reference weights so both models are the same function
        from hydragen_b200.tp import shard_state_dict

        tp_model.load_state_dict(shard_state_dict(full.state_dict(), rank, world), strict=False)
        attn = tp_model.model.layers[0].self_attn
        assert attn.num_heads == 4 // world and attn.num_key_value_heads == 2 // world
        assert attn.q_proj.weight.shape == (4 * 64 // world, 256) and attn.o_proj.weight.shape == (256, 4 * 64 // world)
        ids, logits = _run_generate(tp_model)
        kc = attn.kv_cache.per_completion_k_cache
        assert kc.shape[-2] == 2 // world, kc.shape  # per-rank caches hold local kv heads only
        err = max((a - b).abs().max().item() for a, b in zip(logits, ref_logits))
        same = torch.equal(ids, ref_ids)
        # every rank must have produced the same tokens (replicated sampling, tp.py:178)
        gathered = [torch.zeros_like(ids) for _ in range(world)]
        dist.all_gather(gathered, ids)
        agree = all(torch.equal(gathered[0], x) for x in gathered)
        q.put((rank, err, same, agree, None))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, None, None, None, traceback.format_exc()))


@pytest.mark.timeout(600)
def test_tp2_matches_single_process_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, same, agree, tb in res:
        assert tb is None, f"rank {rank} failed:\n{tb}"
        assert err < 2e-4, f"rank {rank}: tp=2 logits differ from the unsharded model by {err}"
        assert same and agree


def test_shard_state_dict_partitions_heads():
    from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config
    from hydragen_b200.tp import shard_state_dict

    m = HydragenLlamaForCausalLM.from_config(llama_config("tiny"), dtype=torch.float32, device="cpu", seed=1)
    sd = m.state_dict()
    parts = [shard_state_dict(sd, r, 2) for r in range(2)]
    k = "model.layers.0.self_attn.q_proj.weight"
    assert torch.equal(torch.cat([p[k] for p in parts], 0), sd[k])          # column parallel: heads split on dim 0
    k = "model.layers.0.self_attn.o_proj.weight"
    assert torch.equal(torch.cat([p[k] for p in parts], 1), sd[k])          # row parallel
    k = "model.layers.1.mlp.down_proj.weight"
    assert torch.equal(torch.cat([p[k] for p in parts], 1), sd[k])
    assert torch.equal(parts[0]["model.norm.weight"], sd["model.norm.weight"])  # replicated


def test_generate_on_cpu_needs_the_extension():
    """Without the oracle patch the model refuses to run on CPU tensors: no silent fallback."""
    from hydragen_b200._lib import HydragenB200Error
    from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config

    m = HydragenLlamaForCausalLM.from_config(llama_config("tiny"), dtype=torch.float32, device="cpu", seed=1)
    m.setup_caches(2, 4, [1], [8])
    with pytest.raises((HydragenB200Error, ValueError, RuntimeError)):
        m.generate(input_ids=torch.randint(3, 100, (1, 8)), num_return_sequences=2, max_new_tokens=2, temperature=0.0)
