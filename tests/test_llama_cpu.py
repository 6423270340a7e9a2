"""CPU tests of the host logic around the hot path: the Llama surface of
hydragen_b200/llama.py -- cache classes, level accounting of generate(), shared-cache operations, the no-sharing and
single-level baselines, teacher forcing, right-padded (ragged) levels, starting_logits -- with every kernel call routed
to the CPU oracle (tests/oracle_patch.py), in fp32.  On the GPU box tests/test_llama_gpu.py runs the same surface on
the CUDA kernels; here the subject is the Python around them, and the checks are the reference's own self-consistency
criteria (tests/test_e2e.py:122-298 of the reference: Hydragen == no-sharing baseline == single-level baseline), which
in fp32 hold to 1e-4 instead of the reference's fp16 bound of mean rdiff < 0.02."""

import pytest
import torch

import oracle_patch


def _model(**over):
    from hydragen_b200.llama import HydragenLlamaForCausalLM, llama_config

    return HydragenLlamaForCausalLM.from_config(llama_config("tiny", **over), dtype=torch.float32, device="cpu", seed=0, init_std=0.08)


@pytest.fixture()
def model(monkeypatch):
    oracle_patch.apply(monkeypatch)
    torch.manual_seed(0)
    return _model()


def _ids(*shape, seed=0):
    return torch.randint(3, 500, shape, generator=torch.Generator().manual_seed(seed))


def _max_diff(a, b):
    return max((x - y).abs().max().item() for x, y in zip(a, b))


def test_generate_shapes_and_level_accounting(model):
    """num_return_sequences > 1: every id tensor is a shared level, completions form the last level
    (hydragen/llama.py:1232-1308); PRESERVE drops the levels the call added."""
    from hydragen_b200.llama import SharedCacheOp

    model.setup_caches(max_unique_batch_size=6, max_unique_seq_length=5, max_shared_batch_sizes=[1, 2], max_shared_seq_lengths=[20, 7])
    ids = [_ids(1, 20), _ids(2, 7, seed=1)]
    out, logits = model.generate(input_ids=ids, num_return_sequences=3, max_new_tokens=5, temperature=0.0, return_logits=True)
    assert out.shape == (6, 5) and len(logits) == 5 and logits[0].shape == (6, model.config.vocab_size)
    assert model.get_num_used_shared_caches() == 0  # PRESERVE (default)
    out2 = model.generate(input_ids=ids, num_return_sequences=3, max_new_tokens=5, temperature=0.0)
    assert torch.equal(out, out2)  # greedy decoding is deterministic and leaves no state behind
    # EXTEND keeps both levels; a later call can then start from them; WIPE clears first
    model.generate(input_ids=ids, num_return_sequences=3, max_new_tokens=2, temperature=0.0, shared_cache_op=SharedCacheOp.EXTEND)
    assert model.get_num_used_shared_caches() == 2
    assert model.get_shared_cache_len(6).tolist() == [27] * 6
    # WIPE empties the shared levels BEFORE the call and, like EXTEND, keeps what the call adds (hydragen/llama.py:1384-1385
    # truncates only under PRESERVE)
    model.generate(input_ids=ids[0], num_return_sequences=2, max_new_tokens=2, temperature=0.0, shared_cache_op=SharedCacheOp.WIPE)
    assert model.get_num_used_shared_caches() == 1
    assert model.get_shared_cache_len(2).tolist() == [20, 20]


def test_starting_logits_continue_from_cached_levels(model):
    """EXTEND then ``starting_logits``: the second call skips the prefill and must reproduce the first call's tokens
    (hydragen/llama.py:1291-1292)."""
    from hydragen_b200.llama import SharedCacheOp

    model.setup_caches(max_unique_batch_size=4, max_unique_seq_length=6, max_shared_batch_sizes=[1], max_shared_seq_lengths=[16])
    ids = _ids(1, 16, seed=3)
    out, logits = model.generate(input_ids=ids, num_return_sequences=4, max_new_tokens=6, temperature=0.0, return_logits=True,
                                 shared_cache_op=SharedCacheOp.EXTEND)
    assert model.get_num_used_shared_caches() == 1
    again, logits2 = model.generate(starting_logits=logits[0][:1], num_return_sequences=4, max_new_tokens=6, temperature=0.0,
                                    return_logits=True, token_overrides=out)
    assert _max_diff(logits, logits2) < 1e-4
    assert model.get_num_used_shared_caches() == 1  # PRESERVE keeps what was there before the call
    model.empty_shared_cache()
    assert model.get_num_used_shared_caches() == 0


def test_disable_hydragen_and_hierarchy_baselines_agree(model):
    """tests/test_e2e.py:122-298 of the reference: shared-prefix decoding == the no-sharing baseline (prefix copied into
    every sequence's cache) and two shared levels == one shared level + per-sequence suffix."""
    model.setup_caches(max_unique_batch_size=6, max_unique_seq_length=24 + 6, max_shared_batch_sizes=[1, 2], max_shared_seq_lengths=[24, 9])
    one = _ids(1, 24, seed=5)
    out, logits = model.generate(input_ids=one, num_return_sequences=5, max_new_tokens=6, temperature=0.0, return_logits=True)
    _, base = model.generate(input_ids=one, num_return_sequences=5, max_new_tokens=6, temperature=0.0, return_logits=True,
                             token_overrides=out, disable_hydragen=True)
    assert _max_diff(logits, base) < 1e-4
    assert not model.model.get_disable_hydragen()  # the switch is restored
    two = [one, _ids(2, 9, seed=6)]
    out2, logits2 = model.generate(input_ids=two, num_return_sequences=3, max_new_tokens=6, temperature=0.0, return_logits=True)
    _, flat = model.generate(input_ids=two, num_return_sequences=3, max_new_tokens=6, temperature=0.0, return_logits=True,
                             token_overrides=out2, disable_hierarchy=True)
    assert _max_diff(logits2, flat) < 1e-4


def test_unique_suffix_and_teacher_forcing(model):
    """num_return_sequences == 1: the last id tensor is the per-sequence suffix (process_unique); token_overrides
    replace the fed-back tokens, so two runs with different sampling temperature see identical logits."""
    model.setup_caches(max_unique_batch_size=3, max_unique_seq_length=8 + 5, max_shared_batch_sizes=[1], max_shared_seq_lengths=[12])
    ids = [_ids(1, 12, seed=7), _ids(3, 8, seed=8)]
    out, logits = model.generate(input_ids=ids, num_return_sequences=1, max_new_tokens=5, temperature=0.0, return_logits=True)
    assert out.shape == (3, 5)
    torch.manual_seed(1)
    sampled, logits_t = model.generate(input_ids=ids, num_return_sequences=1, max_new_tokens=5, temperature=3.0, return_logits=True,
                                       token_overrides=out)
    assert _max_diff(logits, logits_t) < 1e-5   # same inputs at every step whatever was sampled
    assert sampled.shape == (3, 5)


def test_ragged_level_equals_unpadded_sequences(model):
    """A right-padded shared level with ``seq_lens`` (varlen SharedCache) gives each sequence the logits it gets when its
    own unpadded prompt is run alone."""
    model.setup_caches(max_unique_batch_size=4, max_unique_seq_length=5, max_shared_batch_sizes=[1, 2], max_shared_seq_lengths=[10, 8])
    root = _ids(1, 10, seed=9)
    lvl = _ids(2, 8, seed=10)
    lens = torch.tensor([8, 3])
    out, logits = model.generate(input_ids=[root, lvl], seq_lens=[torch.tensor([10]), lens], num_return_sequences=2, max_new_tokens=4,
                                 temperature=0.0, return_logits=True)
    assert model.model.layers[0].self_attn.kv_cache.shared_caches[1].use_varlen  # the ragged level took the varlen path
    for i, n in enumerate(lens.tolist()):
        solo_ids = [root, lvl[i : i + 1, :n]]
        o_i, l_i = model.generate(input_ids=solo_ids, num_return_sequences=2, max_new_tokens=4, temperature=0.0, return_logits=True,
                                  token_overrides=out[2 * i : 2 * i + 2])
        got = [x[2 * i : 2 * i + 2] for x in logits]
        assert _max_diff(got, l_i) < 1e-4, f"sequence {i} (length {n})"


def test_cache_classes_host_logic():
    """SharedCache.fill packs right-padded rows back to back and records lengths; PerLayerKVCache bookkeeping
    (hydragen/llama.py:58-170, 173-346) -- no kernels involved."""
    from hydragen_b200.llama import PerLayerKVCache, SharedCache

    sc = SharedCache(max_batch_size=3, max_seq_length=6, num_heads=2, head_dim=4, dtype=torch.float32, device=torch.device("cpu"))
    k = torch.arange(3 * 6 * 2 * 4, dtype=torch.float32).view(3, 6, 2, 4)
    lens = torch.tensor([6, 2, 4])
    sc.fill(k, -k, lens)
    assert sc.use_varlen and sc.get_current_batch_size() == 3
    assert sc.get_used_cumsum_lengths().tolist() == [0, 6, 8, 12]
    assert torch.equal(sc.k_cache[6:8], k[1, :2]) and torch.equal(sc.v_cache[8:12], -k[2, :4])
    sc.fill(k[:2], k[:2], torch.tensor([6, 6]))
    assert not sc.use_varlen and sc.sliced_sequence_length == 6 and sc.get_current_batch_size() == 2
    with pytest.raises(ValueError):
        sc.fill(torch.zeros(4, 6, 2, 4), torch.zeros(4, 6, 2, 4), torch.tensor([6] * 4))  # batch exceeds the level
    cache = PerLayerKVCache(4, 8, [1, 2], [6, 5], n_kv_heads=2, head_dim=4, device=torch.device("cpu"), dtype=torch.float32)
    assert not cache.has_shared() and cache.get_shared_len(4).tolist() == [0, 0, 0, 0]
    cache.append_shared(k[:1], k[:1], torch.tensor([6]))
    cache.append_shared(k[:2, :5], k[:2, :5], torch.tensor([5, 3]))
    assert cache.get_shared_len(4).tolist() == [11, 11, 9, 9]  # levels are repeat-interleaved to the batch
    with pytest.raises(ValueError):
        cache.append_shared(k[:1], k[:1], torch.tensor([6]))  # no third level allocated
    cache.truncate_shared_caches(1)
    assert cache.get_shared_len(2).tolist() == [6, 6]
