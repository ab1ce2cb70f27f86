// Direct-space tile kernel, instantiations for KIND_EWALD (see direct_impl.cuh).
#define SNB_DIRECT_KIND KIND_EWALD
#include "direct_impl.cuh"
