// Run-time specialisation of the kernels and LTO link with the user functions.
//
// Reference counterpart: pypde/cfuncs.py:30-75 compiles the user's F/B/S to CPU
// callbacks with numba; here they arrive as device code (LTO-IR from numba's
// CUDA target or nvcc -dlto, PTX, or CUDA source) and are linked *into* the
// hand-written kernels for sm_100a, so every call is inlined.
#pragma once
#include "../../include/pypde_b200.h"
#include <string>
#include <vector>

namespace pypde {

struct KernelConfig {
  int ndim = 1, N = 2, V = 1, flux = 0;
  bool stiff = false, useF = false, useB = false, useS = false, secondOrder = false;
  bool useL = false; // opt-in user wave speed (pypde_b200_set_wavespeed) instead of the eigen-solves
  int dg_cpb = 1;    // cells per block in k_dg
  int faces_fpb = 1; // faces per block in k_faces
  int stiff_wpb = 4; // warps (cells) per block in k_dg_stiff
  int stiff_ks = 6;  // Krylov vectors / Hessenberg columns k_dg_stiff keeps in shared memory
  int stiff_minblocks = 1; // its __launch_bounds__ min-blocks (register budget; set by choose_block_shapes)
  int w3_ti = 4, w3_tj = 4, w3_tk = 8; // k_weno3d tile (PYPDE_B200_W3_TILE=ti,tj,tk)
  bool stiff_stats = false; // PYPDE_B200_STIFF_STATS=1: iteration counters (profiling)
  int ws_block = 512, ws_minblocks = 1; // k_wavespeeds launch bounds (measured best: C5 14.8 ms vs 22.1 at 256 x 2)
  int ff_block = 256, ff_minblocks = 2; // k_faces_fused launch bounds (measured best at C2)
  int fs_block = 256, fs_minblocks = 2; // k_faces_side (experiment)
};

// Chooses the block shapes for a configuration (threads <= 256 where possible,
// shared memory within the 227 KB of an sm_100 CTA).
void choose_block_shapes(KernelConfig &c);

// Returns the linked sm_100a cubin.  Throws std::runtime_error carrying the
// NVRTC / nvJitLink log on failure.  Results are cached in memory and on disk
// ($PYPDE_B200_CACHE, default /tmp/pypde_b200_cache-<uid>).
std::vector<char> build_cubin(const KernelConfig &cfg, const pypde_b200_devfn *F,
                              const pypde_b200_devfn *B, const pypde_b200_devfn *S);

// The opt-in wave-speed function: a process-wide descriptor, copied by
// pypde_b200_set_wavespeed; configurations built while it is set have useL = true and link
//   extern "C" __device__ double user_L(const double *q, const double *dq, int d)
// in place of the finite-difference Jacobian + eigen-solve of max_abs_eigs.
void set_wavespeed(const pypde_b200_devfn *L);      // nullptr clears
bool wavespeed_set();
std::string wavespeed_key();                         // for cache keys (empty when unset)

// The specialised CUDA source (tables + macros + kernels), for diagnostics and
// for the offline nvcc build check.
std::string specialised_source(const KernelConfig &cfg);
std::vector<std::string> specialisation_defines(const KernelConfig &cfg);

} // namespace pypde
