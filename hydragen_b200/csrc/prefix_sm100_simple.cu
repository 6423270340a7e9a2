// Fourth translation unit of the tcgen05 prefix kernel: the non-pipelined ("simple") softmax instantiations,
// compiled in parallel with the others.  See prefix_sm100.cu.
#define HG_PREFIX_TU_SIMPLE 1
#include "prefix_sm100.cu"
