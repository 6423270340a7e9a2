// Second translation unit of the one-CTA-per-unit prefix kernel: the causal (prefill) instantiations, compiled in
// parallel with the unmasked decode-path ones.  See prefix_unit_sm100.cu.
#define HG_UNIT_TU_CAUSAL 1
#include "prefix_unit_sm100.cu"
