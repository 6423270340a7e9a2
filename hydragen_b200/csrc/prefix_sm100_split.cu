// Third translation unit of the tcgen05 prefix kernel: the split-column softmax instantiations (two softmax
// warpgroups per Q tile), compiled in parallel with the others.  See prefix_sm100.cu.
#define HG_PREFIX_TU_SPLIT 1
#include "prefix_sm100.cu"
