/* pypde_b200 — C ABI of the B200-native ADER-WENO stepper.
 *
 * Part 1 is the drop-in boundary: exactly the two entry points of the
 * reference's src/api.h:4-13, with the same argument order, types and
 * semantics (reference caller: pypde/solvers.py:210-212,228-242 through the
 * ctypes argtypes of pypde/utils.py:9-16).  The one difference is what the
 * three function-pointer slots carry: the reference passes CPU callbacks
 * (pypde/cfuncs.py:41-73); here each slot carries a pointer to a
 * `pypde_b200_devfn` descriptor of a *device* function with the same C
 * signature, which the library links into its kernels with nvJitLink for
 * sm_100a.  The slots stay pointer-sized, so ADER_ARGTYPES is unchanged.
 *
 * Part 2 is a handle API over the same solver for callers that keep the state
 * resident in HBM (bench.py, multi-GPU slabs).  Nothing in either part has a
 * CPU fallback: without a CUDA device every compute entry point fails loudly
 * (non-zero return / message on stderr + pypde_b200_last_error()).
 */
#ifndef PYPDE_B200_H
#define PYPDE_B200_H

#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device-function descriptor (replaces the cfunc pointers of cfuncs.py) */
enum {
  PYPDE_B200_LTOIR = 0,       /* NVVM LTO-IR (numba.cuda.compile(..., output='ltoir'), nvcc -dlto) */
  PYPDE_B200_PTX = 1,         /* PTX text, NUL-terminated */
  PYPDE_B200_CUDA_SOURCE = 2  /* CUDA C++ source text, NUL-terminated, compiled with NVRTC */
};

typedef struct pypde_b200_devfn {
  const void *image; /* code image                                          */
  size_t bytes;      /* size of image in bytes (including the NUL for text)  */
  int kind;          /* PYPDE_B200_LTOIR / _PTX / _CUDA_SOURCE               */
  const char *name;  /* label for diagnostics; the image must define the
                        extern "C" __device__ symbol user_F / user_B / user_S:
                          void user_F(double *out, const double *q, const double *dq, int d);
                          void user_B(double *out, const double *q, int d);
                          void user_S(double *out, const double *q);
                        (same argument meaning as reference cfuncs.py:6-8:
                         q[V], dq[ndim][V] row-major, out[V] or out[V][V]) */
} pypde_b200_devfn;

/* ---- Part 1: the reference's C ABI (src/api.h:4-10 and :12-13) ----------- */

/* Replaces reference src/api.h:4-10 / src/api.cpp:5-30 (-> iterator.cpp:38-151).
 * F, B, S point to pypde_b200_devfn descriptors (ignored when useX is false,
 * as api.cpp:21-26).  _u (ncell x V, row-major) is advanced in place to tf;
 * _ret (ndt x ncell*V) receives the snapshots with the reference's row
 * semantics (iterator.cpp:136-139,150).  nThreads is accepted and ignored.
 * Under an initialised multi-GPU communicator (Part 3) _u/_ret are this
 * rank's slab of axis 0 and _nX[0] its local row count. */
void pde_solver(void (*F)(double *, double *, double *, int),
                void (*B)(double *, double *, int), void (*S)(double *, double *),
                bool useF, bool useB, bool useS, double *_u, double tf, int *_nX,
                int ndim, double *_dX, double CFL, int *_boundaryTypes, bool STIFF,
                int FLUX, int N, int V, int ndt, bool secondOrder, double *_ret,
                int nThreads);

/* Replaces reference src/api.h:12-13 / src/api.cpp:32-48: stand-alone WENO
 * reconstruction of an already padded array (m_1..m_n, V) ->
 * (m_1-2(N-1), ..., N, ..., N, V). */
void weno_solver(double *ret, double *_u, int *_nX, int ndim, int N, int V);

/* pde_solver keeps its last solver (kernel module + work arrays in HBM) for the
 * next call with the same configuration; this frees it.  PYPDE_B200_KEEP_SOLVER=0
 * disables the caching altogether. */
int pypde_b200_release_cache(void);

/* ---- Part 2: handle API (state resident in HBM) --------------------------- */
typedef struct pypde_b200_solver pypde_b200_solver;

/* Last error message of the calling thread ("" if none). */
const char *pypde_b200_last_error(void);

/* Library / toolchain probe.  Returns 0 and fills what it can; never needs a GPU. */
int pypde_b200_version(int *nvrtc_major, int *nvrtc_minor, int *nvjitlink_major,
                       int *nvjitlink_minor);

/* Host basis tables for order N (csrc/tables.cpp), for the table parity tests:
 * nodes[N], wghts[N], derv[N*N], endv[2*N], dgmat[N*N], dginv[N*N], sig[N*N],
 * wm[4*N*N] (mL, mR, mCL, mCR), wminv[4*N*N].  Any pointer may be NULL. */
int pypde_b200_tables(int N, double *nodes, double *wghts, double *derv, double *endv,
                      double *dgmat, double *dginv, double *sig, double *wm, double *wminv);

/* The device eigen-solver text (csrc/eig.cuh) compiled for the host, for unit
 * tests of that code: spectral radius of a row-major n x n matrix (n <= 17).
 * qr_only != 0 forces the general QR iteration; *path = 1 if the small-matrix
 * polynomial path decided, 0 if the QR iteration did.  Test aid, not a CPU
 * fallback: no solver entry point calls it. */
int pypde_b200_host_spectral_radius(const double *A, int n, int qr_only, double *rho,
                                    int *path);

/* Same text, the two-matrix entry the fused face kernel uses (n = 3..5): the left and
 * the right state's polynomials searched in lockstep, cold pass then warm pass.
 * rho[2], ok[2]; ok[s] = 0 where the polynomial path deferred to the QR iteration.
 * qr_only modes of the entry above: 1 balanced QR, 2 unbalanced QR, 3 / 4 warm starts
 * from one / three perturbed copies. */
int pypde_b200_host_spectral_radius_pair(const double *A0, const double *A1, int n, double *rho,
                                         int *ok);

/* Same, for the Osher/Roe dissipation y = |A| x = Re(R |Lambda| R^-1 x). */
int pypde_b200_host_abs_matrix_apply(const double *A, int n, const double *x, double *y);

/* Same, the projector form for n = 3..5 the face kernels try first (two simple outer real
 * eigenvalues + one tight cluster, the spectrum of an Euler-type system): 0 = certified and
 * y written, 2 = not certified (the kernel then takes the general routine above). */
int pypde_b200_host_abs_matrix_apply_poly(const double *A, int n, const double *x, double *y);

/* Stand-alone reconstruction with input and output resident in HBM (the device-pointer
 * form of weno_solver, reference api.cpp:32-48): u_dev holds prod(nX) V doubles, ret_dev
 * prod(nX_d - 2(N-1)) N^ndim V doubles; enqueued on `stream` (a CUstream / cudaStream_t,
 * NULL = the legacy default stream) and complete on return. */
int pypde_b200_weno_device(const void *u_dev, void *ret_dev, const int *nX, int ndim, int N, int V,
                           void *stream);

/* OPT-IN, not the reference's numbers: a device function
 *   extern "C" __device__ double user_L(const double *q, const double *dq, int d)
 * returning max |lambda| of dF_d/dQ + B_d at q.  While one is set, solvers created by
 * pde_solver / pypde_b200_create use it in the CFL condition and the Rusanov flux instead of
 * the finite-difference Jacobian + eigen-solve the reference defines (eigs/system.cpp:6-51);
 * results then differ from the reference's at the level of its differencing noise (~1e-8
 * relative in dt).  NULL restores the reference's definition.  The descriptor is copied. */
int pypde_b200_set_wavespeed(const pypde_b200_devfn *L);

/* JIT only (no GPU needed): specialise + link the kernels for a configuration
 * and return the sm_100a cubin size (and optionally the cubin).  Used by the
 * CPU test-suite and by build(). */
int pypde_b200_compile(const pypde_b200_devfn *F, const pypde_b200_devfn *B,
                       const pypde_b200_devfn *S, int ndim, int N, int V, int FLUX,
                       int STIFF, int secondOrder, size_t *cubin_bytes,
                       void *cubin_out, size_t cubin_cap);

/* The specialised CUDA translation unit (#defines + tables + kernels) the JIT
 * compiles for a configuration, as text: for the offline nvcc build check and
 * for reading SASS/PTX.  Copies min(cap, size) bytes; full size in *n. */
int pypde_b200_emit_source(int ndim, int N, int V, int FLUX, int STIFF, int useF, int useB,
                           int useS, int secondOrder, char *out, size_t cap, size_t *n);

/* Create a solver for one slab.  nX/dX/boundaryTypes have ndim entries. */
int pypde_b200_create(pypde_b200_solver **out, const pypde_b200_devfn *F,
                      const pypde_b200_devfn *B, const pypde_b200_devfn *S,
                      const int *nX, int ndim, const double *dX, double CFL,
                      const int *boundaryTypes, int STIFF, int FLUX, int N, int V,
                      int secondOrder);
int pypde_b200_destroy(pypde_b200_solver *s);

/* Run all work of this solver on the given CUstream / cudaStream_t
 * (e.g. torch.cuda.current_stream().cuda_stream).  Default: an own stream. */
int pypde_b200_set_stream(pypde_b200_solver *s, void *stream);

/* State: either copied from/to host memory, or a caller-owned device buffer
 * (ncell x V doubles) that the solver advances in place. */
int pypde_b200_set_state(pypde_b200_solver *s, const double *u_host);
int pypde_b200_get_state(pypde_b200_solver *s, double *u_host);
int pypde_b200_bind_state(pypde_b200_solver *s, void *u_device);

/* Start a run: t = 0, step count = 0, final time tf. */
int pypde_b200_begin(pypde_b200_solver *s, double tf);
/* Enqueue one time step (ghosts, WENO, CFL, dt, predictor, fluxes, update). */
int pypde_b200_step_async(pypde_b200_solver *s);
/* Wait for the enqueued work; returns t, the last dt and the NaN flag. */
int pypde_b200_sync(pypde_b200_solver *s, double *t, double *dt, int *nan_found);
/* Number of kernels this solver has launched so far. */
long long pypde_b200_launch_count(const pypde_b200_solver *s);
/* Per-stage device pointers for stage-wise parity tests:
 * which = 0 ub, 1 w, 2 traces, 3 centers, 4.. face integrals of direction which-4.
 * Copies min(cap, size) doubles to host; returns the full size in *n. */
int pypde_b200_read_stage(pypde_b200_solver *s, int which, double *out, size_t cap,
                          size_t *n);

/* Measurement aids for bench.py.  With profiling on, every kernel launch is
 * bracketed by CUDA events on the launching stream; kernel_times() waits and
 * returns, per kernel name, the summed device milliseconds and launch count
 * since the last call as a text table "name ms launches\n...". */
int pypde_b200_set_profiling(pypde_b200_solver *s, int on);
int pypde_b200_kernel_times(pypde_b200_solver *s, char *out, size_t cap);
/* Measured DFMA throughput of this device, TFLOP/s (FP64 roofline denominator). */
int pypde_b200_fp64_peak(pypde_b200_solver *s, double *tflops);

/* ---- Part 3: multi-GPU slabs (one process per GPU) ------------------------ */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the caller
 * (bench.py uses torch.distributed for the plumbing). */
int pypde_b200_comm_unique_id(void *id128);
int pypde_b200_comm_init(int rank, int nranks, const void *id128);
int pypde_b200_comm_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* PYPDE_B200_H */
