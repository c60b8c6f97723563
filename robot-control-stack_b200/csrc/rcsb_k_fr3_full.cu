// Shape-specialised kernel variant: fr3_empty_world (FR3 + Franka hand, no free bodies), full workspace layout.
#ifndef RCSB_SINGLE_TU
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rcsb.h"
#include "rcsb_layout.h"
#include "rcsb_ctx.cuh"
#include "rcsb_stage.cuh"
#endif
#define RCSB_SHAPE_FR3(maxcon, maxefc, reduced) {9, 9, 8, 9, 24, 182, 1, 1, 1, maxcon, maxefc, 7, 1, 1, 5, reduced, 1, 37, 0}
#define RCSB_VARIANT_NS rcsb_fr3_full
#define RCSB_KERNEL rcsb_k_run_fr3_full
#define RCSB_FIXED_SHAPE RCSB_SHAPE_FR3(16, 58, 0)
#define RCSB_VARIANT_WARPS 9  // what the layout leaves room for: registers per thread follow from it
#include "rcsb_variant.cuh"
#undef RCSB_VARIANT_NS
#undef RCSB_KERNEL
#undef RCSB_FIXED_SHAPE
#undef RCSB_SHAPE_FR3
