// Host driver of the GPU time step (replaces the reference's
// solvers/iterator.cpp:38-151 loop body and its ThreadPool slab scheduler).
#pragma once
#include "dyn.h"
#include "jit.h"
#include <memory>
#include <string>
#include <vector>

namespace pypde {

// must match the structs in kernels.cuh
struct GridParams {
  int nX[3];
  int bt[3];
  double dX[3];
  double rdX[3];
  int halo_lo;
  int halo_hi;
};
struct StepState {
  double t, dt, tf, cfl;
  unsigned long long maxbits;
  long long count;
  int nan_flag;
  int pad;
  unsigned long long stiff_next;
};
struct FluxPtrs {
  CUdeviceptr f[3];
};

// Tile shapes of the TMA-fed WENO kernels (kernels.cuh: W2_* / W3_*); false where the kernel
// is not compiled for the configuration.
struct Weno3dTile {
  int ti, tj, tk;
  size_t smem;
};
bool weno2d_tile(const KernelConfig &c, int *ti_out, int *tj_out);
bool weno3d_tile(const KernelConfig &c, Weno3dTile *t);
// Tensor maps over a padded array `ub` of extents m[] (cells) x V doubles; false if TMA
// cannot address it (odd row pitch) or the driver has no cuTensorMapEncodeTiled.
bool weno2d_map(CUtensorMap *map, CUdeviceptr ub, const long *m, const KernelConfig &c);
bool weno3d_map(CUtensorMap *map, CUdeviceptr ub, const long *m, const KernelConfig &c);

// process-wide slab communicator (one process per GPU)
struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, nranks = 1;
};
Comm &global_comm();

void ensure_context();
void check(CUresult r, const char *what);

struct DeviceBuffer {
  CUdeviceptr p = 0;
  size_t bytes = 0;
  void alloc(size_t n);
  void release();
  ~DeviceBuffer() { release(); }
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
};

class Module {
public:
  Module(const KernelConfig &cfg, const pypde_b200_devfn *F, const pypde_b200_devfn *B,
         const pypde_b200_devfn *S);
  ~Module();
  CUmodule mod = nullptr;
  CUfunction k_fp64_peak = nullptr;
  CUfunction k_boundaries = nullptr, k_weno_sweep = nullptr, k_cfl = nullptr, k_dt = nullptr,
             k_advance = nullptr, k_dg = nullptr, k_faces = nullptr, k_update = nullptr,
             k_wavespeeds = nullptr, k_dg_stiff = nullptr, k_faces_fused = nullptr, k_dg_n = nullptr,
             k_weno2d = nullptr, k_faces_side = nullptr, k_weno3d = nullptr, k_cfl_q = nullptr, k_dg_g = nullptr;
};

class Solver {
public:
  Solver(const KernelConfig &cfg, const pypde_b200_devfn *F, const pypde_b200_devfn *B,
         const pypde_b200_devfn *S, const int *nX, const double *dX, double cfl,
         const int *bt);
  ~Solver();

  void set_stream(CUstream s);
  void set_state(const double *u_host);
  void get_state(double *u_host);
  void bind_state(CUdeviceptr u);
  void begin(double tf);
  void step_async();
  void sync(double *t, double *dt, int *nan_found);
  void snapshot_prev(); // uprev <- u (iterator.cpp:147)
  void get_prev(double *u_host);
  size_t read_stage(int which, double *out, size_t cap);
  // Snapshot pipeline (the `ret` rows of iterator.cpp:136-139): the state is copied
  // device-to-device on the compute stream, then device-to-pinned-host on a copy
  // stream, overlapping the following steps; the pinned buffer is copied into the
  // caller's (pageable) row when its slot is reused or at drain_snapshots().
  void snapshot_async(double *host_row);
  void drain_snapshots();

  // stand-alone reconstruction of an already padded array (api.cpp:32-48)
  static void weno_only(double *ret, const double *u, const int *nX, int ndim, int N, int V);
  // the same with input and output resident in HBM (ret holds prod(nX_d - 2(N-1)) N^ndim V
  // doubles), enqueued on `stream` and complete when this returns
  static void weno_device(CUdeviceptr u, CUdeviceptr ret, const int *nX, int ndim, int N, int V,
                          CUstream stream);

  // per-kernel device times (CUDA events on the launching stream)
  void set_profiling(bool on);
  struct KernelTime {
    std::string name;
    double ms = 0;
    long launches = 0;
  };
  std::vector<KernelTime> kernel_times(); // syncs, returns and clears the record
  // measured DFMA throughput of this GPU in TFLOP/s (micro-kernel in kernels.cuh)
  double measure_fp64_peak();

  long ncell() const { return ncell_; }
  int V() const { return cfg_.V; }
  long long launches = 0;

private:
  void launch(CUfunction f, unsigned grid, unsigned block, size_t smem, void **args,
              const char *name, unsigned grid_y = 1, unsigned grid_z = 1);
  bool profiling_ = false;
  bool ws_pad_set_ = false;
  long ws_pad_ = 0; // PYPDE_B200_WS_SMEM_PAD
  struct Rec {
    const char *name;
    CUevent a, b;
  };
  std::vector<Rec> recs_;
  unsigned grid_for(long total, unsigned block) const;
  void run_sweeps(CUdeviceptr in, const long *shape_in, CUdeviceptr *bufs);
  // Multi-GPU halos (axis 0).  post_halo_exchange() enqueues, on the communication
  // stream, the send / receive of the N edge rows of u per side into halo_lo_/halo_hi_
  // once the compute stream has reached the point where those rows are final; the next
  // k_boundaries waits for it.  Inside a run the exchange is posted right after the edge
  // rows' k_update and overlaps the interior rows' update (and whatever the caller
  // enqueues between steps); halo_valid_ = false forces one at the start of a step.
  // The kernel sequence of one step (all arguments are fixed for the life of the solver:
  // dt, t and the step counter live in the device-side StepState).  On small grids, where a
  // step is ten launches of a few microseconds each, step_async() captures it once into a
  // CUDA graph and replays that (single GPU, profiling off; PYPDE_B200_GRAPH=0/1 overrides
  // the size rule).
  void step_body();
  void drop_graph();
  bool graph_enabled_ = false;
  CUgraphExec graph_exec_ = nullptr;
  long long graph_launches_ = 0;
  void post_halo_exchange();
  void update_cells(long cell0, long ncells);
  CUstream comm_stream_ = nullptr;
  CUevent ev_edge_ = nullptr, ev_halo_ = nullptr;
  bool halo_valid_ = false;

  KernelConfig cfg_;
  std::shared_ptr<Module> mod_;
  GridParams g_;
  double cfl_;
  long ncell_ = 0, ncellw_ = 0, rowlen_ = 0;
  long nfaces_[3] = {0, 0, 0};
  int sms_ = 148;
  CUstream stream_ = nullptr;
  bool own_stream_ = false;

  DeviceBuffer stiff_work_;
  bool fused_faces_ = true;
  bool side_faces_ = true; // two threads per face where k_faces_side exists (PYPDE_B200_FACES_SIDE=0: one)
  bool node_dg_ = true; // k_dg_n / k_dg_g where they apply (PYPDE_B200_DG_NODE=0: always k_dg)
  // 2-D: both WENO sweeps in one tiled kernel fed by TMA (PYPDE_B200_WENO_FUSED=0: two sweeps)
  bool weno2d_ = false;
  int weno2d_ti_ = 0, weno2d_tj_ = 0;
  CUtensorMap ub_map_;
  // 3-D: the three sweeps in one tiled kernel fed by TMA (PYPDE_B200_WENO3D=0: three sweeps)
  bool weno3d_ = false;
  int stiff_wpb_ = 4;
  size_t stiff_smem_ = 0; // dynamic shared memory of k_dg_stiff per block
  long stiff_blocks_ = 0;
  DeviceBuffer u_own_, uprev_, halo_lo_, halo_hi_, ub_, tmpA_, tmpB_, w_, traces_, ws_, centers_,
      flx_[3], state_, qbar_; // qbar_: cell averages of w from k_weno2d for k_cfl_q
  CUdeviceptr u_ = 0;
  StepState *h_state_ = nullptr; // pinned
  struct SnapSlot {
    DeviceBuffer dev;
    double *pinned = nullptr;
    CUevent ready = nullptr, done = nullptr;
    double *dst = nullptr; // pending destination row, or null
  };
  SnapSlot snap_[2];
  int snap_next_ = 0;
  CUstream copy_stream_ = nullptr;
  void finish_slot(SnapSlot &s);
};

} // namespace pypde
