// Force-included before the reference sources (nvcc -include).  The reference
// calls its one-argument checkCUDAError macro with two arguments
// (SfM/sfm.h:13 vs SfM/common.h:18), which MSVC tolerates and gcc does not.
// This re-declares the macro variadically; no reference file is edited.
#pragma once
#include "SfM/common.h"
#undef checkCUDAError
#define checkCUDAError(msg, ...) checkCUDAErrorFn(msg, FILENAME, __LINE__)
