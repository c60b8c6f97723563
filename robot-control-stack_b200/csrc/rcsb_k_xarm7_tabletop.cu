// Shape-specialised kernel variant: xarm7_tabletop (synthetic config C4: xArm7 + table + one free duplo brick), reduced
// workspace layout (the brick resting on the table: 4 contacts as pyramidal cones, 7 friction-loss rows); environments
// whose arm touches the table or the brick are finished by the generic kernel in the full layout.
#ifndef RCSB_SINGLE_TU
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rcsb.h"
#include "rcsb_layout.h"
#include "rcsb_ctx.cuh"
#include "rcsb_stage.cuh"
#endif
#define RCSB_VARIANT_NS rcsb_xarm7_tabletop
#define RCSB_KERNEL rcsb_k_run_xarm7_tabletop
#define RCSB_FIXED_SHAPE {14, 13, 7, 8, 11, 46, 0, 0, 2, 4, 27, 7, 0, 1, 0, 1, 0, 30, 7}
#define RCSB_VARIANT_WARPS 15  // what the layout leaves room for: registers per thread follow from it
#include "rcsb_variant.cuh"
#undef RCSB_VARIANT_NS
#undef RCSB_KERNEL
#undef RCSB_FIXED_SHAPE
