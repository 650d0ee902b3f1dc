/* Test infrastructure, not product code.
 *
 * CPU callbacks (reference cfunc signatures, pypde/cfuncs.py:6-8) compiled by
 * gcc from the SAME C text the GPU side compiles with NVRTC
 * (pypde_b200/systems/systems_src.h), so that the reference library and the
 * CUDA kernels evaluate bit-identical user functions.  -ffp-contract=off.
 */
#include <math.h>
#define PDE_FN

#define SYS_EULER
#define SYS_NDIM 1
#define SYS_F euler_1d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#define SYS_NDIM 2
#define SYS_F euler_2d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#define SYS_NDIM 3
#define SYS_F euler_3d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_EULER

#define SYS_REACTIVE_EULER
#define SYS_NDIM 1
#define SYS_F reactive_euler_1d_F
#define SYS_S reactive_euler_1d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_S
#define SYS_NDIM 2
#define SYS_F reactive_euler_2d_F
#define SYS_S reactive_euler_2d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_S
#undef SYS_REACTIVE_EULER

#define SYS_NAVIER_STOKES
#define SYS_NDIM 1
#define SYS_F navier_stokes_1d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#define SYS_NDIM 2
#define SYS_F navier_stokes_2d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#define SYS_NDIM 3
#define SYS_F navier_stokes_3d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_NAVIER_STOKES

#define SYS_ADVECT_NC
#define SYS_NDIM 1
#define SYS_F advect_nc_1d_F
#define SYS_B advect_nc_1d_B
#define SYS_S advect_nc_1d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_B
#undef SYS_S
#define SYS_NDIM 2
#define SYS_F advect_nc_2d_F
#define SYS_B advect_nc_2d_B
#define SYS_S advect_nc_2d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_B
#undef SYS_S
#define SYS_NDIM 3
#define SYS_F advect_nc_3d_F
#define SYS_B advect_nc_3d_B
#define SYS_S advect_nc_3d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_B
#undef SYS_S
#undef SYS_ADVECT_NC

#define SYS_GPR
#define SYS_NDIM 1
#define SYS_F gpr_1d_F
#define SYS_B gpr_1d_B
#define SYS_S gpr_1d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_B
#undef SYS_S
#define SYS_NDIM 2
#define SYS_F gpr_2d_F
#define SYS_B gpr_2d_B
#define SYS_S gpr_2d_S
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#undef SYS_B
#undef SYS_S
#undef SYS_GPR

#define SYS_BURGERS
#define SYS_NDIM 1
#define SYS_F burgers_1d_F
#include "../pypde_b200/systems/systems_src.h"
#undef SYS_NDIM
#undef SYS_F
#define SYS_NDIM 2
#define SYS_F burgers_2d_F
#include "../pypde_b200/systems/systems_src.h"
