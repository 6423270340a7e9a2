"""hydragen_b200 -- B200-native (sm_100a) implementation of Hydragen's shared-prefix attention
hot path behind the reference's own Python surface.

    from hydragen_b200.attention import hydragen_attention, hydragen_attention_nopad, combine_lse
    from hydragen_b200.flash import flash_attention, flash_attention_varlen, flash_attention_seqlen

Importing the package does not load the CUDA library; the first kernel call does, and raises if
``hydragen_b200/_C/libhydragen_b200.so`` has not been built (``python -m hydragen_b200.build``).
"""

__version__ = "0.1.0"
