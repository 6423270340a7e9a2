"""CPU tests of the checkpoint loaders (hydragen/llama.py:1398-1422 ``from_pretrained``, hydragen/tp.py:135-180
``from_pretrained_tp``): a tiny HuggingFace Llama checkpoint written to disk round-trips into the module tree, full and
head-sharded.  Both loaders build the tree on the ``meta`` device and assign the checkpoint tensors; the RoPE tables
are not module buffers, so nothing of the tree is left on ``meta`` (round-1 advisor finding)."""

import pytest
import torch

from hydragen_b200 import tp as TP
from hydragen_b200.llama import HydragenLlamaForCausalLM


@pytest.fixture(scope="module")
def hf_dir(tmp_path_factory):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                   vocab_size=97, max_position_embeddings=64, rope_theta=10000.0, tie_word_embeddings=False)
    torch.manual_seed(0)
    model = transformers.LlamaForCausalLM(cfg).to(torch.bfloat16)
    d = tmp_path_factory.mktemp("tiny_llama")
    model.save_pretrained(d)
    sd = {k: v.clone() for k, v in model.state_dict().items() if "rotary_emb" not in k}
    return str(d), sd


def _no_meta(model):
    for name, t in list(model.named_parameters()) + list(model.named_buffers()):
        assert t.device.type != "meta", name


def test_from_pretrained_round_trip(hf_dir):
    path, sd = hf_dir
    model = HydragenLlamaForCausalLM.from_pretrained(path, torch_dtype=torch.bfloat16)
    _no_meta(model)
    assert model.dtype == torch.bfloat16
    got = model.state_dict()
    assert set(sd) <= set(got)
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    # the rotary tables exist on a real device, are shared by every layer and are not a submodule of the attention layers
    cos, sin = model.model.rotary_emb.tables(torch.bfloat16, "cpu")
    assert cos.shape == (64, 16) and cos.device.type == "cpu" and torch.isfinite(cos.float()).all()
    assert all(layer.self_attn.rotary_emb is model.model.rotary_emb for layer in model.model.layers)
    assert not any("rotary_emb" in n for n, _ in model.model.layers[0].self_attn.named_modules())


def test_rope_scaling_factor_reaches_the_tables():
    from hydragen_b200.llama import HydragenLlamaModel, llama_config

    base = HydragenLlamaModel(llama_config("tiny")).rotary_emb
    scaled = HydragenLlamaModel(llama_config("tiny", rope_scaling={"type": "linear", "factor": 4.0})).rotary_emb
    assert scaled.scaling_factor == 4.0
    assert torch.allclose(scaled.cos_cached[4], base.cos_cached[1])  # position 4 / factor 4 == position 1


@pytest.mark.parametrize("world", [2])
def test_from_pretrained_tp_round_trip(hf_dir, world, tmp_path, monkeypatch):
    path, sd = hf_dir
    for r in range(world):
        torch.save(TP.shard_state_dict(sd, r, world), tmp_path / f"{r}.pt")
    for r in range(world):
        monkeypatch.setattr(TP, "get_rank", lambda r=r: r)
        monkeypatch.setattr(TP, "get_world_size", lambda: world)
        model = TP.from_pretrained_tp(path, tmp_path, device="cpu")
        _no_meta(model)
        want = TP.shard_state_dict(sd, r, world)
        got = model.state_dict()
        for k, v in want.items():
            assert got[k].shape == v.shape and torch.equal(got[k], v), k
        attn = model.model.layers[0].self_attn
        assert attn.num_heads == 4 // world and attn.num_key_value_heads == 2 // world
        assert attn.q_proj.weight.shape == (64 // world, 64) and attn.o_proj.weight.shape == (64, 64 // world)


def test_generate_capacity_checks():
    """Requests that do not fit the caches or the RoPE table fail on the host before anything is written."""
    from hydragen_b200.llama import llama_config

    model = HydragenLlamaForCausalLM.from_config(llama_config("tiny"), dtype=torch.float32, device="cpu", seed=0)
    model.setup_caches(max_unique_batch_size=4, max_unique_seq_length=16, max_shared_batch_sizes=[1], max_shared_seq_lengths=[32])
    ids = torch.randint(3, 90, (1, 8))
    with pytest.raises(ValueError, match="max_unique_seq_length"):
        model.generate(ids, num_return_sequences=4, max_new_tokens=18)
    with pytest.raises(ValueError, match="max_unique_batch_size"):
        model.generate(ids, num_return_sequences=5, max_new_tokens=4)
    with pytest.raises(ValueError, match="unique cache"):
        model.process_unique(torch.randint(3, 90, (2, 17)))


def test_numa_binding_helper_is_inert_without_a_gpu_and_parses_cpulists(tmp_path, monkeypatch):
    """hydragen_b200.host.bind_process_to_gpu_numa never raises: no CUDA device, no sysfs entry, switched off -> None; the sysfs
    cpulist format is parsed as the kernel documents it."""
    from hydragen_b200.host import bind_process_to_gpu_numa, parse_cpulist

    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == [] and parse_cpulist("5") == [5]
    assert bind_process_to_gpu_numa(0, sysfs_root=str(tmp_path)) is None  # no GPU here / no such PCI device
    monkeypatch.setenv("HYDRAGEN_B200_BIND_NUMA", "0")
    assert bind_process_to_gpu_numa(0) is None


def test_numa_binding_helper_binds_to_the_gpu_local_cpus(tmp_path, monkeypatch):
    """With a (faked) sysfs entry naming a subset of this process's CPUs as local to the GPU, the process is restricted to them."""
    import os
    import types

    import torch

    from hydragen_b200.host import bind_process_to_gpu_numa

    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 2:
        pytest.skip("needs at least two CPUs")
    local = allowed[: len(allowed) // 2]
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    (dev / "local_cpulist").write_text(",".join(str(c) for c in local) + ",100000\n")  # a CPU outside the cpuset is ignored
    (dev / "numa_node").write_text("1\n")
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0))
    try:
        got = bind_process_to_gpu_numa(0, sysfs_root=str(tmp_path))
        assert got == {"pci": "0000:1b:00.0", "numa_node": 1, "cpus": len(local), "of": len(allowed)}
        assert sorted(os.sched_getaffinity(0)) == local
        assert bind_process_to_gpu_numa(0, sysfs_root=str(tmp_path)) is None  # already there: nothing to do
    finally:
        os.sched_setaffinity(0, allowed)
