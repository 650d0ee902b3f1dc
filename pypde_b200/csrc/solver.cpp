#include "solver.h"

#include <map>
#include <mutex>
#include <stdexcept>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

namespace pypde {

namespace {
int ipow(int b, int e) {
  int r = 1;
  while (e-- > 0)
    r *= b;
  return r;
}
} // namespace

// tile of k_weno2d (kernels.cuh: W2_TJ, W2_TI, W2_OK); false where that kernel is not compiled
bool weno2d_tile(const KernelConfig &c, int *ti_out, int *tj_out) {
  if (c.ndim != 2)
    return false;
  const int H = 2 * (c.N - 1), N = c.N, V = c.V;
  const int tj = V * (32 + H) <= 256 ? 32 : (V * (16 + H) <= 256 ? 16 : 8);
  const int row = (tj + H) * V;
  auto smem = [&](int ti) { return ((ti + H) * row + ti * (tj + H) * N * V) * 8 + 256; };
  const int ti = smem(8) <= 48 * 1024 ? 8 : 4;
  if (row > 256 || smem(ti) > 48 * 1024)
    return false;
  if (ti_out)
    *ti_out = ti;
  if (tj_out)
    *tj_out = tj;
  return true;
}

// tile of k_weno3d (kernels.cuh: W3_TI, W3_TJ, W3_TK, W3_SMEM, W3_OK)
bool weno3d_tile(const KernelConfig &c, Weno3dTile *t) {
  if (c.ndim != 3)
    return false;
  const long H = 2 * (c.N - 1), N = c.N, V = c.V;
  const long ti = c.w3_ti, tj = c.w3_tj, tk = (c.w3_tk + H) * V <= 256 ? c.w3_tk : 4;
  const long row = (tk + H) * V;
  const long in = (ti + H) * (tj + H) * row, a = ti * (tj + H) * (tk + H) * N * V,
             b = ti * tj * (tk + H) * N * N * V;
  const long smem = ((in > b ? in : b) + a) * 8 + 128;
  if (row > 256 || smem > 200 * 1024)
    return false;
  if (t)
    *t = Weno3dTile{(int)ti, (int)tj, (int)tk, (size_t)smem};
  return true;
}

bool weno2d_map(CUtensorMap *map, CUdeviceptr ub, const long *m, const KernelConfig &c) {
  int ti, tj;
  const DriverApi &d = driver();
  if (!weno2d_tile(c, &ti, &tj) || !d.TensorMapEncodeTiled)
    return false;
  const cuuint64_t cols = (cuuint64_t)m[1] * c.V, rows = (cuuint64_t)m[0];
  const int H = 2 * (c.N - 1);
  // (row pitch and base address must be multiples of 16 bytes)
  if (cols % 2 != 0 || (ub & 15) != 0)
    return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)((tj + H) * c.V), (cuuint32_t)(ti + H)};
  const cuuint32_t estr[2] = {1, 1};
  return d.TensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ub, dims, strides,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ub as a 3-D tensor [m0][m1][m2 V doubles]
bool weno3d_map(CUtensorMap *map, CUdeviceptr ub, const long *m, const KernelConfig &c) {
  Weno3dTile t;
  const DriverApi &d = driver();
  if (!weno3d_tile(c, &t) || !d.TensorMapEncodeTiled)
    return false;
  const cuuint64_t cols = (cuuint64_t)m[2] * c.V;
  const int H = 2 * (c.N - 1);
  if (cols % 2 != 0 || (ub & 15) != 0)
    return false;
  const cuuint64_t dims[3] = {cols, (cuuint64_t)m[1], (cuuint64_t)m[0]};
  const cuuint64_t strides[2] = {cols * sizeof(double), cols * (cuuint64_t)m[1] * sizeof(double)};
  const cuuint32_t box[3] = {(cuuint32_t)((t.tk + H) * c.V), (cuuint32_t)(t.tj + H),
                             (cuuint32_t)(t.ti + H)};
  const cuuint32_t estr[3] = {1, 1, 1};
  return d.TensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)ub, dims, strides,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// kernels.cuh: DGN_OK — where k_dg_n is compiled
bool dgn_ok(const KernelConfig &c) {
  const int Nd = ipow(c.N, c.ndim);
  if (c.useB || c.secondOrder || c.N < 2 || Nd > 32)
    return false;
  const int cpw = 32 / Nd, fs = c.ndim * c.N * c.V * Nd;
  const int sm_cell = fs + ((Nd % 16) + 16 - (fs % 16)) % 16;
  return (long)cpw * sm_cell * 8 <= 48 * 1024;
}

// kernels.cuh: DGG_OK — where k_dg_g is compiled
bool dgg_ok(const KernelConfig &c) {
  const int Nd = ipow(c.N, c.ndim);
  if (!(c.useB || c.secondOrder) || c.N < 2 || Nd > 32)
    return false;
  const int cpw = 32 / Nd, qs = c.N * c.V * Nd, raw = 2 * qs + c.ndim * qs;
  const int sm_cell = raw + ((Nd % 16) + 16 - (raw % 16)) % 16;
  return (long)cpw * sm_cell * 8 <= 48 * 1024;
}

void check(CUresult r, const char *what) {
  if (r == CUDA_SUCCESS)
    return;
  const char *s = nullptr;
  driver().GetErrorString(r, &s);
  throw std::runtime_error(std::string("pypde_b200: ") + what + ": " + (s ? s : "unknown CUDA error"));
}

void ensure_context() {
  const DriverApi &d = driver();
  CUcontext ctx = nullptr;
  check(d.CtxGetCurrent(&ctx), "cuCtxGetCurrent");
  if (ctx)
    return;
  int dev = 0;
  const char *e = getenv("PYPDE_B200_DEVICE");
  if (!e || !*e)
    e = getenv("LOCAL_RANK");
  if (e && *e)
    dev = atoi(e);
  int n = 0;
  check(d.DeviceGetCount(&n), "cuDeviceGetCount");
  if (n == 0)
    throw std::runtime_error("pypde_b200: no CUDA device visible (this library has no CPU path)");
  if (dev >= n)
    dev = dev % n;
  CUdevice cd;
  check(d.DeviceGet(&cd, dev), "cuDeviceGet");
  check(d.DevicePrimaryCtxRetain(&ctx, cd), "cuDevicePrimaryCtxRetain");
  check(d.CtxSetCurrent(ctx), "cuCtxSetCurrent");
}

Comm &global_comm() {
  static Comm c;
  return c;
}

void DeviceBuffer::alloc(size_t n) {
  release();
  if (n == 0)
    n = 8;
  check(driver().MemAlloc(&p, n), "cuMemAlloc");
  bytes = n;
}
void DeviceBuffer::release() {
  if (p) {
    driver().MemFree(p);
    p = 0;
    bytes = 0;
  }
}

// ---------------------------------------------------------------------------
Module::Module(const KernelConfig &cfg, const pypde_b200_devfn *F, const pypde_b200_devfn *B,
               const pypde_b200_devfn *S) {
  std::vector<char> cubin = build_cubin(cfg, F, B, S);
  ensure_context();
  const DriverApi &d = driver();
  check(d.ModuleLoadData(&mod, cubin.data()), "cuModuleLoadData (is this GPU sm_100?)");
  auto get = [&](CUfunction &f, const char *name) {
    check(d.ModuleGetFunction(&f, mod, name), name);
  };
  get(k_boundaries, "k_boundaries");
  get(k_weno_sweep, "k_weno_sweep");
  get(k_cfl, "k_cfl");
  get(k_dt, "k_dt");
  get(k_advance, "k_advance");
  get(k_dg, "k_dg");
  get(k_faces, "k_faces");
  get(k_update, "k_update");
  get(k_fp64_peak, "k_fp64_peak");
  if (cfg.useF)
    get(k_wavespeeds, "k_wavespeeds");
  if (cfg.stiff)
    get(k_dg_stiff, "k_dg_stiff");
  if (cfg.useF && cfg.flux == 0)
    get(k_faces_fused, "k_faces_fused");
  if (dgn_ok(cfg))
    get(k_dg_n, "k_dg_n");
  if (dgg_ok(cfg))
    get(k_dg_g, "k_dg_g");
  if (weno2d_tile(cfg, nullptr, nullptr))
    get(k_weno2d, "k_weno2d");
  if (!cfg.useB && !cfg.secondOrder) // kernels.cuh: !NEED_GRAD
    get(k_cfl_q, "k_cfl_q");
  {
    Weno3dTile t3;
    if (weno3d_tile(cfg, &t3)) {
      get(k_weno3d, "k_weno3d");
      if (t3.smem > 48 * 1024)
        check(d.FuncSetAttribute(k_weno3d, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                 (int)t3.smem),
              "cuFuncSetAttribute(k_weno3d smem)");
    }
  }
  if (cfg.useF && cfg.flux == 0 && !cfg.useB)
    get(k_faces_side, "k_faces_side");
}

Module::~Module() {
  if (mod)
    driver().ModuleUnload(mod);
}

// ---------------------------------------------------------------------------
Solver::Solver(const KernelConfig &cfg, const pypde_b200_devfn *F, const pypde_b200_devfn *B,
               const pypde_b200_devfn *S, const int *nX, const double *dX, double cfl,
               const int *bt)
    : cfg_(cfg), cfl_(cfl) {
  if (cfg_.ndim < 1 || cfg_.ndim > 3)
    throw std::runtime_error("pypde_b200: ndim must be 1, 2 or 3");
  if (cfg_.V < 1)
    throw std::runtime_error("pypde_b200: V must be >= 1");
  if (cfg_.N < 1)
    throw std::runtime_error("pypde_b200: order N must be >= 1");
  if (cfg_.flux < 0 || cfg_.flux > 2)
    throw std::runtime_error("pypde_b200: FLUX must be 0 (rusanov), 1 (roe) or 2 (osher)");
  choose_block_shapes(cfg_);
  // k_faces_fused (both sides of a face point in one thread) wins where the eigen-solves
  // stay in registers: small systems with a first-order flux.  Measured on B200: C2 (V = 4)
  // 6.2 against 9.3 ms per step; GPR (V = 17) 19.9 against 13.4 ms and 3-D Navier-Stokes
  // (second-order flux: gradient traces of both sides live at once) 20.5 against 12.9 ms
  // for k_wavespeeds + k_faces.
  // Round 2: with TWO threads per face (k_faces_side) the second-order case wins too — a
  // lane holds one side's state and gradient only: 3-D Navier-Stokes at 32 x 128^2,
  // k_wavespeeds + k_faces 18.4 + 2.6 ms against 15.8 ms (profiles/r2_c5_sweep.txt), and one
  // pass over the 58 GB of traces instead of two.
  if (const char *e = getenv("PYPDE_B200_WS_SMEM_PAD"))
    ws_pad_ = atol(e);
  fused_faces_ = cfg_.V <= 5 && !(cfg_.secondOrder && cfg_.useB);
  if (const char *e = getenv("PYPDE_B200_FUSED_FACES"))
    fused_faces_ = *e != '0';
  // two threads per face (k_faces_side) where that kernel exists: measured 5.37 against
  // 6.02 ms per step for k_faces_fused at C2, same bits
  side_faces_ = true;
  if (const char *e = getenv("PYPDE_B200_FACES_SIDE"))
    side_faces_ = *e != '0';
  if (const char *e = getenv("PYPDE_B200_DG_NODE"))
    node_dg_ = *e != '0';
  {
    long cells = 1;
    for (int i = 0; i < cfg_.ndim; i++)
      cells *= nX[i];
    graph_enabled_ = cells <= 65536; // beyond that a step is milliseconds of kernels
    if (const char *e = getenv("PYPDE_B200_GRAPH"))
      graph_enabled_ = *e != '0';
  }
  // PYPDE_B200_PREBUILD=1 (build box, no GPU): JIT this configuration's cubin into the
  // on-disk cache before the device is asked for, so that a later run on a GPU box that
  // shares the cache directory ($PYPDE_B200_CACHE) starts without compiling
  if (const char *e = getenv("PYPDE_B200_PREBUILD"))
    if (*e == '1')
      build_cubin(cfg_, F, B, S);
  ensure_context();
  const DriverApi &d = driver();
  CUdevice dev;
  check(d.CtxGetDevice(&dev), "cuCtxGetDevice");
  d.DeviceGetAttribute(&sms_, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev);
  if (sms_ <= 0)
    sms_ = 148;

  mod_ = std::make_shared<Module>(cfg_, F, B, S);

  const int nd = cfg_.ndim, N = cfg_.N, V = cfg_.V;
  memset(&g_, 0, sizeof g_);
  for (int i = 0; i < 3; i++) {
    g_.nX[i] = 1;
    g_.dX[i] = 1.;
    g_.rdX[i] = 1.;
  }
  ncell_ = 1;
  ncellw_ = 1;
  long ncellb = 1;
  for (int i = 0; i < nd; i++) {
    if (nX[i] < 1)
      throw std::runtime_error("pypde_b200: empty grid axis");
    g_.nX[i] = nX[i];
    g_.dX[i] = dX[i];
    g_.rdX[i] = 1. / dX[i];
    g_.bt[i] = bt[i];
    ncell_ *= nX[i];
    ncellw_ *= nX[i] + 2;
    ncellb *= nX[i] + 2 * N;
  }
  rowlen_ = ncell_ / nX[0];

  const Comm &cm = global_comm();
  if (cm.nranks > 1) {
    if (nX[0] < N)
      throw std::runtime_error("pypde_b200: a slab needs at least N rows along axis 0");
    const bool periodic = bt[0] == 1;
    g_.halo_lo = (cm.rank > 0 || periodic) ? 2 : 0;
    g_.halo_hi = (cm.rank < cm.nranks - 1 || periodic) ? 2 : 0;
  } else {
    g_.halo_lo = g_.halo_hi = bt[0] == 1 ? 1 : 0;
  }

  const int Nd = ipow(N, nd);
  const int NP = N * ipow(N, nd - 1);
  const int TRW = 1 + (cfg_.secondOrder ? nd : 0);
  const size_t D = sizeof(double);
  halo_lo_.alloc((size_t)N * rowlen_ * V * D);
  halo_hi_.alloc((size_t)N * rowlen_ * V * D);
  ub_.alloc((size_t)ncellb * V * D);
  // WENO intermediates
  {
    long shape[3];
    for (int i = 0; i < nd; i++)
      shape[i] = nX[i] + 2 * N;
    size_t sizes[3] = {0, 0, 0};
    for (int dd = 0; dd < nd; dd++) {
      shape[dd] -= 2 * (N - 1);
      size_t n = (size_t)ipow(N, dd + 1) * V;
      for (int i = 0; i < nd; i++)
        n *= shape[i];
      sizes[dd] = n;
    }
    if (nd >= 2)
      tmpA_.alloc(sizes[0] * D);
    if (nd >= 3)
      tmpB_.alloc(sizes[1] * D);
    w_.alloc((size_t)ncellw_ * Nd * V * D);
  }
  traces_.alloc((size_t)ncellw_ * 2 * nd * NP * TRW * V * D);
  // k_weno2d / k_weno3d: TMA descriptor of ub as a 2-D tensor [nX_0+2N rows][(nX_1+2N) V
  // doubles] / 3-D tensor (row pitch must be a multiple of 16 bytes; otherwise the
  // sweep-by-sweep path stays)
  {
    long mb[3] = {1, 1, 1};
    for (int i = 0; i < nd; i++)
      mb[i] = nX[i] + 2 * N;
    if (nd == 2 && weno2d_tile(cfg_, &weno2d_ti_, &weno2d_tj_) && mod_->k_weno2d) {
      const char *e = getenv("PYPDE_B200_WENO_FUSED");
      weno2d_ = !(e && *e == '0') && weno2d_map(&ub_map_, ub_.p, mb, cfg_);
      // the tile kernel also leaves the cell averages the CFL condition needs (k_cfl_q)
      const char *q = getenv("PYPDE_B200_CFL_Q");
      if (weno2d_ && mod_->k_cfl_q && !(q && *q == '0'))
        qbar_.alloc((size_t)ncellw_ * V * D);
    }
    // OPT-IN (PYPDE_B200_WENO3D=1).  Measured on B200 (profiles/r2_weno3d.txt; 3-D
    // Navier-Stokes, N = 3, 32 x 128 x 128): three k_weno_sweep launches 0.43 ms, k_weno3d
    // 0.47-0.61 ms depending on the tile.  The reconstruction is bound by the FP64 pipe,
    // not by HBM (63 % pipe utilisation in either form), and a tile recomputes the first two
    // sweeps on its halo: 1.27 x the arithmetic at the 4 x 4 x 8 tile that fits shared
    // memory, executed at the same rate.  In 2-D the halo is cheaper and the tile kernel
    // wins (0.61 against 0.71 ms at 2048^2); in 3-D the sweeps stay the default.
    if (nd == 3 && mod_->k_weno3d) {
      const char *e = getenv("PYPDE_B200_WENO3D");
      weno3d_ = e && *e == '1' && weno3d_map(&ub_map_, ub_.p, mb, cfg_);
    }
  }
  // per trace point: lambda, [lambda_visc] — only where the two-kernel face path runs
  if (cfg_.useF && !(cfg_.flux == 0 && fused_faces_))
    ws_.alloc((size_t)ncellw_ * 2 * nd * NP * (1 + (cfg_.secondOrder ? 1 : 0)) * D);
  centers_.alloc((size_t)ncellw_ * V * D);
  const int FLXW = cfg_.useB ? 2 : 1;
  for (int dd = 0; dd < nd; dd++) {
    long nf = 1;
    for (int i = 0; i < nd; i++)
      nf *= nX[i] + (i == dd ? 1 : 0);
    nfaces_[dd] = nf;
    flx_[dd].alloc((size_t)nf * FLXW * V * D);
  }
  state_.alloc(sizeof(StepState));
  check(d.MemAllocHost((void **)&h_state_, sizeof(StepState)), "cuMemAllocHost");
  memset(h_state_, 0, sizeof(StepState));

  check(d.StreamCreate(&stream_, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
  own_stream_ = true;

  if (cfg_.stiff) {
    // one warp per cell, cells drawn from a device-side queue; global workspace per warp
    // (kernels.cuh: NK_WORK) for the Krylov vectors beyond the shared-memory resident ones
    const size_t n = (size_t)N * Nd * V;
    const size_t nk_work = 41 * n + 10 * n + 42 * 41 + 6 * 41;
    stiff_wpb_ = cfg_.stiff_wpb; // = PDE_STIFF_WPB of the compiled kernel
    const size_t smem_warp =
        ((3 + nd + cfg_.stiff_ks) * n + (size_t)32 * 35 / 2 + 5 * 41 + 2) * D;
    if (smem_warp > 220 * 1024)
      throw std::runtime_error("pypde_b200: stiff predictor working set exceeds shared memory");
    stiff_smem_ = stiff_wpb_ * smem_warp;
    if (stiff_smem_ > 48 * 1024)
      check(d.FuncSetAttribute(mod_->k_dg_stiff, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                               (int)stiff_smem_),
            "cuFuncSetAttribute(k_dg_stiff smem)");
    // As many blocks as are resident at once, no more: the cells come from a queue, and the
    // global workspace (one slice per warp of the grid) then stays within L2 — ncu on a
    // grid of twice that size: 22 GB of workspace written through to DRAM per launch.
    long blocks = (ncellw_ + stiff_wpb_ - 1) / stiff_wpb_;
    int per_sm = 0;
    if (d.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mod_->k_dg_stiff, 32 * stiff_wpb_,
                                                    stiff_smem_) != CUDA_SUCCESS ||
        per_sm < 1)
      per_sm = 4;
    long cap = (long)sms_ * per_sm;
    stiff_blocks_ = blocks < cap ? blocks : cap;
    // (+ 64 bytes: the iteration counters of PDE_STIFF_STATS)
    stiff_work_.alloc((size_t)stiff_blocks_ * stiff_wpb_ * nk_work * D + 64);
    check(d.MemsetD8Async(stiff_work_.p + (size_t)stiff_blocks_ * stiff_wpb_ * nk_work * D, 0, 64,
                          stream_),
          "cuMemsetD8Async(stiff stats)");
  }
  // dynamic shared memory opt-in
  const size_t dg_smem = (size_t)cfg_.dg_cpb * (2 + nd) * N * Nd * V * D;
  if (dg_smem > 48 * 1024)
    check(d.FuncSetAttribute(mod_->k_dg, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                             (int)dg_smem),
          "cuFuncSetAttribute(k_dg smem)");
}

void Solver::finish_slot(SnapSlot &s) {
  if (!s.dst)
    return;
  check(driver().EventSynchronize(s.done), "cuEventSynchronize(snapshot)");
  memcpy(s.dst, s.pinned, (size_t)ncell_ * cfg_.V * sizeof(double));
  s.dst = nullptr;
}

void Solver::snapshot_async(double *host_row) {
  ensure_context();
  const DriverApi &d = driver();
  const size_t bytes = (size_t)ncell_ * cfg_.V * sizeof(double);
  if (!copy_stream_)
    check(d.StreamCreate(&copy_stream_, CU_STREAM_NON_BLOCKING), "cuStreamCreate(copy)");
  SnapSlot &s = snap_[snap_next_];
  snap_next_ ^= 1;
  finish_slot(s); // the slot's previous snapshot must have landed in its row
  if (!s.pinned) {
    s.dev.alloc(bytes);
    check(d.MemAllocHost((void **)&s.pinned, bytes), "cuMemAllocHost(snapshot)");
    check(d.EventCreate(&s.ready, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    check(d.EventCreate(&s.done, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
  }
  check(d.MemcpyDtoDAsync(s.dev.p, u_, bytes, stream_), "cuMemcpyDtoDAsync(snapshot)");
  check(d.EventRecord(s.ready, stream_), "cuEventRecord");
  check(d.StreamWaitEvent(copy_stream_, s.ready, 0), "cuStreamWaitEvent");
  check(d.MemcpyDtoHAsync(s.pinned, s.dev.p, bytes, copy_stream_), "cuMemcpyDtoHAsync(snapshot)");
  check(d.EventRecord(s.done, copy_stream_), "cuEventRecord");
  s.dst = host_row;
}

void Solver::drain_snapshots() {
  for (SnapSlot &s : snap_)
    finish_slot(s);
}

Solver::~Solver() {
  const DriverApi &d = driver();
  if (stream_)
    d.StreamSynchronize(stream_);
  drop_graph();
  for (SnapSlot &s : snap_) {
    if (s.dst && s.done)
      d.EventSynchronize(s.done);
    if (s.pinned)
      d.MemFreeHost(s.pinned);
    if (s.ready)
      d.EventDestroy(s.ready);
    if (s.done)
      d.EventDestroy(s.done);
  }
  if (copy_stream_) {
    d.StreamSynchronize(copy_stream_);
    d.StreamDestroy(copy_stream_);
  }
  if (comm_stream_) {
    d.StreamSynchronize(comm_stream_);
    d.StreamDestroy(comm_stream_);
    d.EventDestroy(ev_edge_);
    d.EventDestroy(ev_halo_);
  }
  if (stream_)
    d.StreamSynchronize(stream_);
  if (own_stream_ && stream_)
    d.StreamDestroy(stream_);
  if (h_state_)
    d.MemFreeHost(h_state_);
}

void Solver::set_stream(CUstream s) {
  const DriverApi &d = driver();
  if (stream_)
    check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
  if (comm_stream_)
    check(d.StreamSynchronize(comm_stream_), "cuStreamSynchronize(comm)");
  if (own_stream_ && stream_)
    d.StreamDestroy(stream_);
  stream_ = s;
  own_stream_ = false;
  drop_graph();
  // the legacy default stream cannot be captured: plain launches there
  if (s == nullptr || s == (CUstream)0x1 /* CU_STREAM_LEGACY */)
    graph_enabled_ = false;
}

void Solver::set_state(const double *u_host) {
  ensure_context();
  if (!u_) {
    u_own_.alloc((size_t)ncell_ * cfg_.V * sizeof(double));
    u_ = u_own_.p;
    drop_graph();
  }
  const DriverApi &d = driver();
  check(d.MemcpyHtoDAsync(u_, u_host, (size_t)ncell_ * cfg_.V * sizeof(double), stream_),
        "cuMemcpyHtoDAsync(u)");
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
  halo_valid_ = false;
}

void Solver::get_state(double *u_host) {
  ensure_context();
  if (!u_)
    throw std::runtime_error("pypde_b200: no state set");
  const DriverApi &d = driver();
  check(d.MemcpyDtoHAsync(u_host, u_, (size_t)ncell_ * cfg_.V * sizeof(double), stream_),
        "cuMemcpyDtoHAsync(u)");
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
}

void Solver::bind_state(CUdeviceptr u) {
  u_own_.release();
  u_ = u;
  halo_valid_ = false;
  drop_graph(); // the graph holds the state pointer
}

void Solver::snapshot_prev() {
  if (!uprev_.p)
    uprev_.alloc((size_t)ncell_ * cfg_.V * sizeof(double));
  check(driver().MemcpyDtoDAsync(uprev_.p, u_, (size_t)ncell_ * cfg_.V * sizeof(double), stream_),
        "cuMemcpyDtoDAsync(uprev)");
}

void Solver::get_prev(double *u_host) {
  if (!uprev_.p)
    throw std::runtime_error("pypde_b200: no previous state kept");
  const DriverApi &d = driver();
  check(d.MemcpyDtoHAsync(u_host, uprev_.p, (size_t)ncell_ * cfg_.V * sizeof(double), stream_),
        "cuMemcpyDtoHAsync(uprev)");
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
}

void Solver::begin(double tf) {
  ensure_context();
  if (!u_)
    throw std::runtime_error("pypde_b200: set or bind a state before begin()");
  StepState s;
  memset(&s, 0, sizeof s);
  s.t = 0.;
  s.dt = 0.;
  s.tf = tf;
  s.cfl = cfl_;
  *h_state_ = s;
  halo_valid_ = false; // the caller may have written the (bound) state since the last step
  const DriverApi &d = driver();
  check(d.MemcpyHtoDAsync(state_.p, h_state_, sizeof(StepState), stream_), "cuMemcpyHtoDAsync(state)");
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
}

unsigned Solver::grid_for(long total, unsigned block) const {
  long blocks = (total + block - 1) / block;
  long cap = (long)sms_ * 16;
  if (blocks > cap)
    blocks = cap;
  if (blocks < 1)
    blocks = 1;
  return (unsigned)blocks;
}

void Solver::launch(CUfunction f, unsigned grid, unsigned block, size_t smem, void **args,
                    const char *name, unsigned grid_y, unsigned grid_z) {
  const DriverApi &d = driver();
  Rec r{name, nullptr, nullptr};
  if (profiling_) {
    check(d.EventCreate(&r.a, CU_EVENT_DEFAULT), "cuEventCreate");
    check(d.EventCreate(&r.b, CU_EVENT_DEFAULT), "cuEventCreate");
    check(d.EventRecord(r.a, stream_), "cuEventRecord");
  }
  check(d.LaunchKernel(f, grid, grid_y, grid_z, block, 1, 1, (unsigned)smem, stream_, args, nullptr),
        name);
  if (profiling_) {
    check(d.EventRecord(r.b, stream_), "cuEventRecord");
    recs_.push_back(r);
  }
  launches++;
}

void Solver::set_profiling(bool on) { profiling_ = on; }

std::vector<Solver::KernelTime> Solver::kernel_times() {
  ensure_context();
  const DriverApi &d = driver();
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
  std::vector<KernelTime> out;
  for (Rec &r : recs_) {
    float ms = 0;
    check(d.EventElapsedTime(&ms, r.a, r.b), "cuEventElapsedTime");
    d.EventDestroy(r.a);
    d.EventDestroy(r.b);
    KernelTime *kt = nullptr;
    for (KernelTime &k : out)
      if (k.name == r.name)
        kt = &k;
    if (!kt) {
      out.push_back(KernelTime{r.name, 0., 0});
      kt = &out.back();
    }
    kt->ms += ms;
    kt->launches++;
  }
  recs_.clear();
  return out;
}

double Solver::measure_fp64_peak() {
  ensure_context();
  const DriverApi &d = driver();
  DeviceBuffer sink;
  sink.alloc(sizeof(double) * 1024);
  int iters = 4096;
  const unsigned block = 256, grid = (unsigned)sms_ * 8;
  double seed = 1e-9;
  void *args[] = {&sink.p, &iters, &seed};
  CUevent a, b;
  check(d.EventCreate(&a, CU_EVENT_DEFAULT), "cuEventCreate");
  check(d.EventCreate(&b, CU_EVENT_DEFAULT), "cuEventCreate");
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    check(d.EventRecord(a, stream_), "cuEventRecord");
    check(d.LaunchKernel(mod_->k_fp64_peak, grid, 1, 1, block, 1, 1, 0, stream_, args, nullptr),
          "k_fp64_peak");
    check(d.EventRecord(b, stream_), "cuEventRecord");
    check(d.EventSynchronize(b), "cuEventSynchronize");
    float ms = 0;
    check(d.EventElapsedTime(&ms, a, b), "cuEventElapsedTime");
    // 16 independent chains x iters FMAs x 2 flop per thread
    double flops = (double)grid * block * 16.0 * iters * 2.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best)
      best = tf;
  }
  d.EventDestroy(a);
  d.EventDestroy(b);
  return best;
}

void Solver::post_halo_exchange() {
  const Comm &cm = global_comm();
  if (cm.nranks <= 1)
    return;
  const DriverApi &d = driver();
  if (!comm_stream_) {
    check(d.StreamCreate(&comm_stream_, CU_STREAM_NON_BLOCKING), "cuStreamCreate(comm)");
    check(d.EventCreate(&ev_edge_, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
    check(d.EventCreate(&ev_halo_, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
  }
  // the edge rows are final at this point of the compute stream
  check(d.EventRecord(ev_edge_, stream_), "cuEventRecord(edge)");
  check(d.StreamWaitEvent(comm_stream_, ev_edge_, 0), "cuStreamWaitEvent(edge)");
  const NcclApi &nc = nccl();
  const int N = cfg_.N, V = cfg_.V;
  const size_t cnt = (size_t)N * rowlen_ * V;
  const bool periodic = g_.bt[0] == 1;
  const int lo = cm.rank > 0 ? cm.rank - 1 : (periodic ? cm.nranks - 1 : -1);
  const int hi = cm.rank < cm.nranks - 1 ? cm.rank + 1 : (periodic ? 0 : -1);
  const CUdeviceptr first_rows = u_;
  const CUdeviceptr last_rows = u_ + ((size_t)(g_.nX[0] - N) * rowlen_ * V) * sizeof(double);
  auto ok = [&](int r, const char *what) {
    if (r != 0)
      throw std::runtime_error(std::string("pypde_b200: NCCL ") + what + ": " + nc.GetErrorString(r));
  };
  // Sends are posted [first rows -> low, last rows -> high] and receives
  // [high halo <- high, low halo <- low]: with two ranks on a periodic axis both
  // neighbours are the same peer and NCCL matches messages between a pair in
  // posting order, so the peer's "first rows" must meet our high-halo receive.
  ok(nc.GroupStart(), "ncclGroupStart");
  if (lo >= 0)
    ok(nc.Send((const void *)first_rows, cnt, NCCL_FLOAT64, lo, cm.comm, comm_stream_), "ncclSend");
  if (hi >= 0)
    ok(nc.Send((const void *)last_rows, cnt, NCCL_FLOAT64, hi, cm.comm, comm_stream_), "ncclSend");
  if (hi >= 0)
    ok(nc.Recv((void *)halo_hi_.p, cnt, NCCL_FLOAT64, hi, cm.comm, comm_stream_), "ncclRecv");
  if (lo >= 0)
    ok(nc.Recv((void *)halo_lo_.p, cnt, NCCL_FLOAT64, lo, cm.comm, comm_stream_), "ncclRecv");
  ok(nc.GroupEnd(), "ncclGroupEnd");
  check(d.EventRecord(ev_halo_, comm_stream_), "cuEventRecord(halo)");
  halo_valid_ = true;
}

void Solver::update_cells(long cell0, long ncells) {
  if (ncells <= 0)
    return;
  FluxPtrs fp;
  for (int i = 0; i < 3; i++)
    fp.f[i] = flx_[i < cfg_.ndim ? i : 0].p;
  void *args[] = {&u_, &centers_.p, &fp, &g_, &state_.p, &cell0, &ncells};
  launch(mod_->k_update, grid_for(ncells * cfg_.V, 256), 256, 0, args, "k_update");
}

// runs the ndim sweeps; shape_in = padded shape of `in`; bufs[d] = output of sweep d
void Solver::run_sweeps(CUdeviceptr in, const long *shape_in, CUdeviceptr *bufs) {
  const int nd = cfg_.ndim, N = cfg_.N, V = cfg_.V;
  long shape[3];
  for (int i = 0; i < nd; i++)
    shape[i] = shape_in[i];
  CUdeviceptr cur = in;
  for (int dd = 0; dd < nd; dd++) {
    long n1 = 1, n3 = 1;
    for (int i = 0; i < dd; i++)
      n1 *= shape[i];
    for (int i = dd + 1; i < nd; i++)
      n3 *= shape[i];
    int md = (int)shape[dd];
    long n34 = n3 * ipow(N, dd);
    int n1i = (int)n1;
    long total = n1 * (md - 2 * (N - 1)) * n34 * V;
    CUdeviceptr out = bufs[dd];
    void *args[] = {&cur, &out, &n1i, &md, &n34};
    launch(mod_->k_weno_sweep, grid_for(total, 256), 256, 0, args, "k_weno_sweep");
    cur = out;
    shape[dd] -= 2 * (N - 1);
  }
}

void Solver::drop_graph() {
  if (graph_exec_) {
    driver().GraphExecDestroy(graph_exec_);
    graph_exec_ = nullptr;
  }
}

void Solver::step_async() {
  ensure_context();
  if (!(graph_enabled_ && !profiling_ && global_comm().nranks <= 1)) {
    step_body();
    return;
  }
  const DriverApi &d = driver();
  if (!graph_exec_) {
    // capture one step (nothing executes yet), instantiate, then replay below
    const long long l0 = launches;
    CUgraph graph = nullptr;
    check(d.StreamBeginCapture(stream_, CU_STREAM_CAPTURE_MODE_RELAXED), "cuStreamBeginCapture");
    try {
      step_body();
    } catch (...) {
      d.StreamEndCapture(stream_, &graph);
      if (graph)
        d.GraphDestroy(graph);
      throw;
    }
    check(d.StreamEndCapture(stream_, &graph), "cuStreamEndCapture");
    graph_launches_ = launches - l0;
    launches = l0;
    CUresult r = d.GraphInstantiate(&graph_exec_, graph, 0);
    d.GraphDestroy(graph);
    check(r, "cuGraphInstantiate");
  }
  check(d.GraphLaunch(graph_exec_, stream_), "cuGraphLaunch");
  launches += graph_launches_;
}

void Solver::step_body() {
  const int nd = cfg_.ndim, N = cfg_.N, V = cfg_.V;
  const int Nd = ipow(N, nd);
  const Comm &cm = global_comm();

  if (cm.nranks > 1) {
    // halos of the current state: posted at the end of the previous step (overlapped
    // with its interior update), or here if the state was set since
    if (!halo_valid_)
      post_halo_exchange();
    check(driver().StreamWaitEvent(stream_, ev_halo_, 0), "cuStreamWaitEvent(halo)");
    halo_valid_ = false;
  }
  {
    long total = 1;
    for (int i = 0; i < nd; i++)
      total *= g_.nX[i] + 2 * N;
    total *= V;
    void *args[] = {&u_, &halo_lo_.p, &halo_hi_.p, &ub_.p, &g_};
    launch(mod_->k_boundaries, grid_for(total, 256), 256, 0, args, "k_boundaries");
  }
  if (weno2d_) {
    int n0 = g_.nX[0] + 2, n1 = g_.nX[1] + 2;
    void *args[] = {&ub_map_, &w_.p, &n0, &n1, &qbar_.p};
    launch(mod_->k_weno2d, (unsigned)((n0 + weno2d_ti_ - 1) / weno2d_ti_), 256, 0, args, "k_weno2d",
           (unsigned)((n1 + weno2d_tj_ - 1) / weno2d_tj_));
  } else if (weno3d_) {
    Weno3dTile t3;
    weno3d_tile(cfg_, &t3);
    int n0 = g_.nX[0] + 2, n1 = g_.nX[1] + 2, n2 = g_.nX[2] + 2;
    void *args[] = {&ub_map_, &w_.p, &n0, &n1, &n2};
    launch(mod_->k_weno3d, (unsigned)((n0 + t3.ti - 1) / t3.ti), 256, t3.smem, args, "k_weno3d",
           (unsigned)((n1 + t3.tj - 1) / t3.tj), (unsigned)((n2 + t3.tk - 1) / t3.tk));
  } else {
    long shape[3];
    for (int i = 0; i < nd; i++)
      shape[i] = g_.nX[i] + 2 * N;
    CUdeviceptr bufs[3];
    if (nd == 1)
      bufs[0] = w_.p;
    else if (nd == 2) {
      bufs[0] = tmpA_.p;
      bufs[1] = w_.p;
    } else {
      bufs[0] = tmpA_.p;
      bufs[1] = tmpB_.p;
      bufs[2] = w_.p;
    }
    run_sweeps(ub_.p, shape, bufs);
  }
  if (qbar_.p) {
    void *args[] = {&qbar_.p, &ncellw_, &g_, &state_.p};
    launch(mod_->k_cfl_q, grid_for(ncellw_, 128), 128, 0, args, "k_cfl_q");
  } else {
    void *args[] = {&w_.p, &ncellw_, &g_, &state_.p};
    launch(mod_->k_cfl, grid_for(ncellw_, 128), 128, 0, args, "k_cfl");
  }
  if (cm.nranks > 1) {
    const NcclApi &nc = nccl();
    CUdeviceptr mb = state_.p + offsetof(StepState, maxbits);
    int r = nc.AllReduce((const void *)mb, (void *)mb, 1, NCCL_UINT64, NCCL_MAX, cm.comm, stream_);
    if (r != 0)
      throw std::runtime_error(std::string("pypde_b200: ncclAllReduce: ") + nc.GetErrorString(r));
  }
  {
    void *args[] = {&state_.p};
    launch(mod_->k_dt, 1, 1, 0, args, "k_dt");
  }
  if (cfg_.stiff) {
    void *args[] = {&w_.p, &traces_.p, &centers_.p, &ncellw_, &g_, &state_.p, &stiff_work_.p};
    launch(mod_->k_dg_stiff, (unsigned)stiff_blocks_, 32 * stiff_wpb_, stiff_smem_, args,
           "k_dg_stiff");
  } else if (node_dg_ && mod_->k_dg_n) {
    // one thread per spatial node, 32 / N^ndim cells per one-warp block
    const int cpw = 32 / Nd;
    long nblocks = (ncellw_ + cpw - 1) / cpw;
    long cap = (long)sms_ * 32;
    if (nblocks > cap)
      nblocks = cap;
    void *args[] = {&w_.p, &traces_.p, &centers_.p, &ncellw_, &g_, &state_.p};
    launch(mod_->k_dg_n, (unsigned)nblocks, (unsigned)(cpw * Nd), 0, args, "k_dg_n");
  } else if (node_dg_ && mod_->k_dg_g) {
    // the same with gradient terms (B, second-order flux)
    const int cpw = 32 / Nd;
    long nblocks = (ncellw_ + cpw - 1) / cpw;
    long cap = (long)sms_ * 32;
    if (nblocks > cap)
      nblocks = cap;
    void *args[] = {&w_.p, &traces_.p, &centers_.p, &ncellw_, &g_, &state_.p};
    launch(mod_->k_dg_g, (unsigned)nblocks, (unsigned)(cpw * Nd), 0, args, "k_dg_g");
  } else {
    const unsigned block = cfg_.dg_cpb * N * Nd;
    const size_t smem = (size_t)cfg_.dg_cpb * (2 + nd) * N * Nd * V * sizeof(double);
    long nblocks = (ncellw_ + cfg_.dg_cpb - 1) / cfg_.dg_cpb;
    long cap = (long)sms_ * 32;
    if (nblocks > cap)
      nblocks = cap;
    void *args[] = {&w_.p, &traces_.p, &centers_.p, &ncellw_, &g_, &state_.p};
    launch(mod_->k_dg, (unsigned)nblocks, block, smem, args, "k_dg");
  }
  // Rusanov: one fused pass per direction (k_faces_fused); PYPDE_B200_FUSED_FACES=0 keeps
  // the two-kernel path (k_wavespeeds + k_faces), which Roe / Osher / no-F always use
  if (cfg_.useF && cfg_.flux == 0 && fused_faces_) {
    for (int dd = 0; dd < nd; dd++) {
      void *args[] = {&traces_.p, &flx_[dd].p, &dd, &nfaces_[dd], &g_, &state_.p};
      if (side_faces_ && mod_->k_faces_side)
        launch(mod_->k_faces_side, grid_for(2 * nfaces_[dd], cfg_.fs_block), cfg_.fs_block, 0, args,
               "k_faces_side");
      else
        launch(mod_->k_faces_fused, grid_for(nfaces_[dd], cfg_.ff_block), cfg_.ff_block, 0, args,
               "k_faces_fused");
    }
  } else if (cfg_.useF && cfg_.flux == 0) {
    long total = ncellw_ * 2 * nd;
    void *args[] = {&traces_.p, &ws_.p, &ncellw_, &g_};
    // (experiment knob: unused dynamic shared memory caps the resident blocks per SM — the
    //  n > 5 eigen-solves keep their matrices in local memory, and how many of them are in
    //  flight decides whether that working set stays in L2; tools/variant_sweep.py eig)
    const long pad = ws_pad_;
    if (pad > 48 * 1024 && !ws_pad_set_) {
      check(driver().FuncSetAttribute(mod_->k_wavespeeds,
                                      CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)pad),
            "cuFuncSetAttribute(k_wavespeeds smem pad)");
      ws_pad_set_ = true;
    }
    launch(mod_->k_wavespeeds, grid_for(total, cfg_.ws_block), cfg_.ws_block, (size_t)pad, args,
           "k_wavespeeds");
  }
  if ((cfg_.useF || cfg_.useB) && !(cfg_.useF && cfg_.flux == 0 && fused_faces_)) {
    const int NP = N * ipow(N, nd - 1);
    const int FLXW = cfg_.useB ? 2 : 1;
    for (int dd = 0; dd < nd; dd++) {
      const unsigned block = cfg_.faces_fpb * NP;
      const size_t smem = (size_t)cfg_.faces_fpb * NP * FLXW * V * sizeof(double);
      long nblocks = (nfaces_[dd] + cfg_.faces_fpb - 1) / cfg_.faces_fpb;
      long cap = (long)sms_ * 32;
      if (nblocks > cap)
        nblocks = cap;
      void *args[] = {&traces_.p, &ws_.p, &flx_[dd].p, &dd, &nfaces_[dd], &g_, &state_.p};
      launch(mod_->k_faces, (unsigned)nblocks, block, smem, args, "k_faces");
    }
  }
  if (cm.nranks > 1 && g_.nX[0] >= 2 * N) {
    // the N rows each neighbour needs first; their exchange then overlaps the rest
    const long edge = (long)N * rowlen_;
    update_cells(0, edge);
    update_cells(ncell_ - edge, edge);
    post_halo_exchange();
    update_cells(edge, ncell_ - 2 * edge);
  } else {
    update_cells(0, ncell_);
    if (cm.nranks > 1)
      post_halo_exchange();
  }
  if (cm.nranks > 1) {
    // The NaN stop of iterator.cpp:141-145 must be the same decision on every rank:
    // k_update raises nan_flag from this rank's own cells only, and a rank that left the
    // time loop alone would leave the others waiting in the next dt reduction / halo
    // exchange for ever.  One 4-byte max-all-reduce per step (~10 us against milliseconds).
    const NcclApi &nc = nccl();
    CUdeviceptr nf = state_.p + offsetof(StepState, nan_flag);
    int r = nc.AllReduce((const void *)nf, (void *)nf, 1, NCCL_INT32, NCCL_MAX, cm.comm, stream_);
    if (r != 0)
      throw std::runtime_error(std::string("pypde_b200: ncclAllReduce(nan): ") + nc.GetErrorString(r));
  }
  {
    void *args[] = {&state_.p};
    launch(mod_->k_advance, 1, 1, 0, args, "k_advance");
  }
}

void Solver::sync(double *t, double *dt, int *nan_found) {
  ensure_context();
  const DriverApi &d = driver();
  check(d.MemcpyDtoHAsync(h_state_, state_.p, sizeof(StepState), stream_), "cuMemcpyDtoHAsync(state)");
  check(d.StreamSynchronize(stream_), "cuStreamSynchronize (a kernel failed?)");
  if (t)
    *t = h_state_->t;
  if (dt)
    *dt = h_state_->dt;
  if (nan_found)
    *nan_found = h_state_->nan_flag;
}

size_t Solver::read_stage(int which, double *out, size_t cap) {
  ensure_context();
  const DeviceBuffer *b = nullptr;
  switch (which) {
  case 0:
    b = &ub_;
    break;
  case 1:
    b = &w_;
    break;
  case 2:
    b = &traces_;
    break;
  case 3:
    b = &centers_;
    break;
  case 7:
    b = &ws_;
    break;
  case 8: { // k_dg_stiff iteration counters (PYPDE_B200_STIFF_STATS=1): 4 x u64 as raw doubles
    if (!stiff_work_.p)
      throw std::runtime_error("pypde_b200: not a stiff solver");
    const DriverApi &dd = driver();
    if (out && cap >= 4) {
      check(dd.MemcpyDtoHAsync(out, stiff_work_.p + stiff_work_.bytes - 64, 32, stream_),
            "cuMemcpyDtoHAsync(stiff stats)");
      check(dd.StreamSynchronize(stream_), "cuStreamSynchronize");
    }
    return 4;
  }
  default:
    if (which >= 4 && which < 4 + cfg_.ndim)
      b = &flx_[which - 4];
  }
  if (!b || !b->p)
    throw std::runtime_error("pypde_b200: unknown stage");
  size_t n = b->bytes / sizeof(double);
  size_t m = n < cap ? n : cap;
  const DriverApi &d = driver();
  if (m && out) {
    check(d.MemcpyDtoHAsync(out, b->p, m * sizeof(double), stream_), "cuMemcpyDtoHAsync(stage)");
    check(d.StreamSynchronize(stream_), "cuStreamSynchronize");
  }
  return n;
}

// ---------------------------------------------------------------------------
void Solver::weno_device(CUdeviceptr u, CUdeviceptr ret, const int *nX, int ndim, int N, int V,
                         CUstream st) {
  if (ndim < 1 || ndim > 3)
    throw std::runtime_error("pypde_b200: weno_solver supports ndim 1..3");
  KernelConfig cfg;
  cfg.ndim = ndim;
  cfg.N = N;
  cfg.V = V;
  choose_block_shapes(cfg);
  if (const char *e = getenv("PYPDE_B200_PREBUILD"))
    if (*e == '1')
      build_cubin(cfg, nullptr, nullptr, nullptr);
  ensure_context();
  const DriverApi &d = driver();
  // the kernel module of a (ndim, N, V) reconstruction is kept for the next call
  static std::mutex mods_mutex;
  static std::map<std::string, std::shared_ptr<Module>> mods;
  std::shared_ptr<Module> modp;
  {
    char key[96];
    CUcontext ctx = nullptr;
    d.CtxGetCurrent(&ctx);
    snprintf(key, sizeof key, "%d,%d,%d,%d,%d,%d,%p", ndim, N, V, cfg.w3_ti, cfg.w3_tj, cfg.w3_tk,
             (void *)ctx);
    std::lock_guard<std::mutex> lk(mods_mutex);
    std::shared_ptr<Module> &slot = mods[key];
    if (!slot)
      slot = std::make_shared<Module>(cfg, nullptr, nullptr, nullptr);
    modp = slot;
  }
  Module &mod = *modp;
  int sms = 148;
  long shape[3] = {1, 1, 1};
  for (int i = 0; i < ndim; i++) {
    shape[i] = nX[i];
    if (nX[i] < 2 * N - 1)
      throw std::runtime_error("pypde_b200: weno_solver needs at least 2N-1 cells per axis");
  }
  // the TMA-fed tile kernels (all sweeps in one pass) where they apply, as in the time step
  CUtensorMap map;
  const char *e2 = getenv("PYPDE_B200_WENO_FUSED"), *e3 = getenv("PYPDE_B200_WENO3D");
  if (ndim == 2 && mod.k_weno2d && !(e2 && *e2 == '0') && weno2d_map(&map, u, shape, cfg)) {
    int ti, tj;
    weno2d_tile(cfg, &ti, &tj);
    int n0 = (int)shape[0] - 2 * (N - 1), n1 = (int)shape[1] - 2 * (N - 1);
    CUdeviceptr no_qbar = 0;
    void *args[] = {&map, &ret, &n0, &n1, &no_qbar};
    check(d.LaunchKernel(mod.k_weno2d, (unsigned)((n0 + ti - 1) / ti), (unsigned)((n1 + tj - 1) / tj),
                         1, 256, 1, 1, 0, st, args, nullptr),
          "cuLaunchKernel(k_weno2d)");
    check(d.StreamSynchronize(st), "cuStreamSynchronize");
    return;
  }
  if (ndim == 3 && mod.k_weno3d && e3 && *e3 == '1' && weno3d_map(&map, u, shape, cfg)) {
    Weno3dTile t3;
    weno3d_tile(cfg, &t3);
    int n0 = (int)shape[0] - 2 * (N - 1), n1 = (int)shape[1] - 2 * (N - 1),
        n2 = (int)shape[2] - 2 * (N - 1);
    void *args[] = {&map, &ret, &n0, &n1, &n2};
    check(d.LaunchKernel(mod.k_weno3d, (unsigned)((n0 + t3.ti - 1) / t3.ti),
                         (unsigned)((n1 + t3.tj - 1) / t3.tj), (unsigned)((n2 + t3.tk - 1) / t3.tk),
                         256, 1, 1, (unsigned)t3.smem, st, args, nullptr),
          "cuLaunchKernel(k_weno3d)");
    check(d.StreamSynchronize(st), "cuStreamSynchronize");
    return;
  }
  DeviceBuffer buf[2];
  CUdeviceptr cur = u;
  for (int dd = 0; dd < ndim; dd++) {
    long n1 = 1, n3 = 1;
    for (int i = 0; i < dd; i++)
      n1 *= shape[i];
    for (int i = dd + 1; i < ndim; i++)
      n3 *= shape[i];
    int md = (int)shape[dd];
    long n34 = n3 * ipow(N, dd);
    int n1i = (int)n1;
    long total = n1 * (md - 2 * (N - 1)) * n34 * V;
    CUdeviceptr out = ret; // the last sweep writes the result
    if (dd < ndim - 1) {
      DeviceBuffer &o = buf[dd & 1];
      o.alloc((size_t)total * N * sizeof(double));
      out = o.p;
    }
    void *args[] = {&cur, &out, &n1i, &md, &n34};
    long blocks = (total + 255) / 256;
    if (blocks > sms * 16)
      blocks = sms * 16;
    if (blocks < 1)
      blocks = 1;
    check(d.LaunchKernel(mod.k_weno_sweep, (unsigned)blocks, 1, 1, 256, 1, 1, 0, st, args, nullptr),
          "cuLaunchKernel(k_weno_sweep)");
    cur = out;
    shape[dd] -= 2 * (N - 1);
  }
  // (the intermediates are released on return)
  check(d.StreamSynchronize(st), "cuStreamSynchronize");
}

void Solver::weno_only(double *ret, const double *u, const int *nX, int ndim, int N, int V) {
  if (ndim < 1 || ndim > 3)
    throw std::runtime_error("pypde_b200: weno_solver supports ndim 1..3");
  ensure_context();
  const DriverApi &d = driver();
  size_t nin = V, nout = V;
  for (int i = 0; i < ndim; i++) {
    if (nX[i] < 2 * N - 1)
      throw std::runtime_error("pypde_b200: weno_solver needs at least 2N-1 cells per axis");
    nin *= nX[i];
    nout *= (size_t)(nX[i] - 2 * (N - 1)) * N;
  }
  CUstream st;
  check(d.StreamCreate(&st, CU_STREAM_NON_BLOCKING), "cuStreamCreate");
  DeviceBuffer in, out;
  in.alloc(nin * sizeof(double));
  out.alloc(nout * sizeof(double));
  check(d.MemcpyHtoDAsync(in.p, u, nin * sizeof(double), st), "cuMemcpyHtoDAsync");
  weno_device(in.p, out.p, nX, ndim, N, V, st);
  check(d.MemcpyDtoHAsync(ret, out.p, nout * sizeof(double), st), "cuMemcpyDtoHAsync");
  check(d.StreamSynchronize(st), "cuStreamSynchronize");
  d.StreamDestroy(st);
}

} // namespace pypde
