"""Case list and input checksum shared by make_golden.py (build container only) and the tests
(which must not touch /root/reference)."""

import torch

from oracle import hydragen_oracle as O


def case_list():
    """(name, sizes, qheads, kvheads, dim, dtype, seed, nq)"""
    cases = []
    # tests/test_attention.py:17-32 : fp16, qheads 8, kvheads {1, 8}, d 128
    for ci, sizes in enumerate(O.REFERENCE_SIZES_LIST):
        for kvh in (1, 8):
            cases.append((f"ref{ci}_kv{kvh}_fp16", sizes, 8, kvh, 128, "float16", 0, 1))
    # BASELINE.json cfg#1: B=4, one prefix len 32, suffix 8, 4 heads d=64, fp32
    cases.append(("cfg1_toy_fp32", [[32], [8, 8, 8, 8]], 4, 4, 64, "float32", 0, 1))
    cases.append(("cfg1_toy_ragged_fp32", [[32], [8, 3, 1, 5]], 4, 4, 64, "float32", 1, 1))
    # beyond the reference's coverage (last paragraph)
    cases.append(("bf16_gqa4_d128", [[40], [5, 9, 2, 7, 1, 3]], 8, 2, 128, "bfloat16", 2, 1))
    cases.append(("bf16_d64", [[70, 70], [4, 4, 4, 4]], 4, 4, 64, "bfloat16", 3, 1))
    cases.append(("fp16_nq3_causal", [[33], [5, 5]], 8, 8, 128, "float16", 4, 3))
    cases.append(("bf16_3level_varlen", [[300], [17, 130], [3, 1, 20, 64]], 8, 4, 128, "bfloat16", 5, 1))
    cases.append(("bf16_b256_prefix512", [[512], list(range(1, 257))], 4, 4, 128, "bfloat16", 6, 1))
    return cases


DT = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}


def checksum(case):
    s = case["q"].double().sum() + case["k"].double().sum() * 3 + case["v"].double().sum() * 7
    for a, b in zip(case["shared_ks"], case["shared_vs"]):
        s = s + a.double().sum() * 11 + b.double().sum() * 13
    return float(s)
