// Second translation unit of the tcgen05 prefix kernel: the causal (prefill) instantiations, compiled in
// parallel with the unmasked decode-path ones.  See prefix_sm100.cu.
#define HG_PREFIX_TU_CAUSAL 1
#include "prefix_sm100.cu"
