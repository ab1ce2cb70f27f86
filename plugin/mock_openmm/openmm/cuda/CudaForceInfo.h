/* STAND-IN (compile check only): see CudaContext.h */
#include "CudaContext.h"
