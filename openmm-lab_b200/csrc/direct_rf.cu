// Direct-space tile kernel, instantiations for KIND_RF (see direct_impl.cuh).
#define SNB_DIRECT_KIND KIND_RF
#include "direct_impl.cuh"
