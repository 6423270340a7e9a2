// Fifth translation unit of the tcgen05 prefix kernel: the alternate-block softmax instantiations (two softmax
// warpgroups per Q tile taking the key blocks in turn), compiled in parallel with the others.  See prefix_sm100.cu.
// Status: compiled, not yet run on hardware (round-2 first experiment); never selected unless
// HYDRAGEN_B200_PREFIX_SOFTMAX=alt.
#define HG_PREFIX_TU_ALT 1
#include "prefix_sm100.cu"
