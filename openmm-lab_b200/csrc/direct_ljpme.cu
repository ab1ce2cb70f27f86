// Direct-space tile kernel, instantiations for KIND_LJPME (see direct_impl.cuh).
#define SNB_DIRECT_KIND KIND_LJPME
#include "direct_impl.cuh"
