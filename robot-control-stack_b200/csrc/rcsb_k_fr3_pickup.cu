// Shape-specialised kernel variant: fr3_simple_pick_up (FR3 + Franka hand + one free box), reduced workspace layout
// (the cube resting on the floor: 4 contacts; a grasp is finished by the generic kernel in the full layout).
#ifndef RCSB_SINGLE_TU
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rcsb.h"
#include "rcsb_layout.h"
#include "rcsb_ctx.cuh"
#include "rcsb_stage.cuh"
#endif
#define RCSB_VARIANT_NS rcsb_fr3_pickup
#define RCSB_KERNEL rcsb_k_run_fr3_pickup
#define RCSB_FIXED_SHAPE {16, 15, 8, 10, 25, 206, 1, 1, 2, 4, 17, 7, 1, 1, 5, 1, 1, 47, 0}
#define RCSB_VARIANT_WARPS 12  // what the layout leaves room for: registers per thread follow from it
#include "rcsb_variant.cuh"
#undef RCSB_VARIANT_NS
#undef RCSB_KERNEL
#undef RCSB_FIXED_SHAPE
