// kernels_base.cuh -- includes, launch macros and integer types shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tfbs.h"
#include "tables.hpp"

// Launch syntax and the dynamic shared-memory declaration are spelled as macros: tests/cuda_emu redefines them to run these very
// kernels, thread by thread, on the host of the GPU-less build container (a test of the kernel logic; not a product path).
#ifndef TFBS_LAUNCH
#define TFBS_LAUNCH(kernel, grid, block, smem, stream) kernel<<<(grid), (block), (smem), (stream)>>>
#endif
#ifndef TFBS_DYNAMIC_SHARED
#define TFBS_DYNAMIC_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace tfbs {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

}  // namespace tfbs
