// The reference's drivers #include this translation unit of its unity build (e.g. examples/Compression_thermal.cpp:22-26).
// In weldform_b200 the code it held runs on the GPU behind the C ABI (include/weldform_b200.h); the host-side member
// functions of SPH::Domain are defined inline in Domain.h. This file exists so that those drivers compile unchanged.
#include "Domain.h"
