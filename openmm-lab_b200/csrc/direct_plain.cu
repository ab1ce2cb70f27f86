// Direct-space tile kernel, instantiations for KIND_PLAIN (see direct_impl.cuh).
#define SNB_DIRECT_KIND KIND_PLAIN
#include "direct_impl.cuh"
