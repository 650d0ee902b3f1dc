#include "jit.h"
#include "dyn.h"
#include "tables.h"

#include <map>
#include <stdint.h>
#include <mutex>
#include <stdexcept>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <limits.h>
#include <sys/stat.h>
#include <unistd.h>

namespace pypde {

// kernels.cuh, embedded at build time (see __graft_entry__.build / Makefile)
static const char *const KERNEL_SOURCE =
#include "kernels_embed.inc"
    ;

namespace {

int ipow(int b, int e) {
  int r = 1;
  while (e-- > 0)
    r *= b;
  return r;
}

uint64_t fnv1a(const void *data, size_t n, uint64_t h) {
  const unsigned char *p = (const unsigned char *)data;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

struct Hasher {
  uint64_t a = 14695981039346656037ull, b = 0x9e3779b97f4a7c15ull;
  void add(const void *d, size_t n) {
    a = fnv1a(d, n, a);
    b = fnv1a(d, n, b ^ (n * 0x100000001b3ull));
  }
  void add(const std::string &s) { add(s.data(), s.size()); }
  std::string hex() const {
    char buf[40];
    snprintf(buf, sizeof buf, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
    return buf;
  }
};

// On-disk cubin cache.  The cubins in it are loaded as trusted GPU code, so the directory
// must be this user's alone: default $XDG_CACHE_HOME/pypde_b200 or ~/.cache/pypde_b200
// (PYPDE_B200_CACHE overrides), created 0700, and refused — the disk cache is then simply
// not used — unless lstat shows a real directory (no symlink) owned by getuid() that
// neither group nor others can write or read.  Returns "" when there is no usable one.
std::string cache_dir() {
  std::string d;
  const char *e = getenv("PYPDE_B200_CACHE");
  if (e && *e)
    d = e;
  else {
    const char *x = getenv("XDG_CACHE_HOME");
    const char *h = getenv("HOME");
    if (x && *x == '/')
      d = std::string(x);
    else if (h && *h == '/') {
      d = std::string(h) + "/.cache";
      mkdir(d.c_str(), 0700); // (may exist with the user's own mode: only our leaf is checked)
    } else
      return std::string();
    d += "/pypde_b200";
  }
  mkdir(d.c_str(), 0700);
  struct stat st;
  if (lstat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid() ||
      (st.st_mode & 077) != 0)
    return std::string();
  return d;
}

// Cubins shipped next to the library: <directory of libpypde.so>/cubin_cache, filled by
// __graft_entry__.build() / tools/prebuild_cache.sh (PYPDE_B200_CACHE pointed there) for the
// configurations of the tests and benches.  Read-only at run time, looked up after the
// user's cache, and trusted under the same rule: a real directory owned by the calling user
// that neither group nor others can write.  Returns "" when there is none.
std::string shipped_cache_dir() {
  Dl_info info;
  if (!dladdr((void *)&shipped_cache_dir, &info) || !info.dli_fname)
    return std::string();
  // (the library's real location: a symlink to it elsewhere — the reference package's
  //  pypde/build/libpypde.so of INTEGRATION.md 1 — still finds the cubins shipped with it)
  char real[PATH_MAX];
  std::string d(realpath(info.dli_fname, real) ? real : info.dli_fname);
  const size_t slash = d.rfind('/');
  if (slash == std::string::npos)
    return std::string();
  d = d.substr(0, slash) + "/cubin_cache";
  struct stat st;
  const int rc = lstat(d.c_str(), &st);
  if (getenv("PYPDE_B200_CACHE_DEBUG"))
    fprintf(stderr, "pypde_b200: library '%s', shipped cache '%s': lstat %d, mode %o, uid %d (caller %d)\n",
            info.dli_fname, d.c_str(), rc, rc == 0 ? (unsigned)st.st_mode : 0u,
            rc == 0 ? (int)st.st_uid : -1, (int)getuid());
  if (rc != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid() || (st.st_mode & 022) != 0)
    return std::string();
  return d;
}

bool read_file(const std::string &path, std::vector<char> &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f)
    return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? n : 0);
  bool ok = n > 0 && fread(out.data(), 1, n, f) == (size_t)n;
  fclose(f);
  return ok;
}

void write_file_atomic(const std::string &path, const std::vector<char> &data) {
  char tmp[512];
  snprintf(tmp, sizeof tmp, "%s.%d.tmp", path.c_str(), (int)getpid());
  FILE *f = fopen(tmp, "wb");
  if (!f)
    return;
  bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (ok)
    rename(tmp, path.c_str());
  else
    unlink(tmp);
}

// PYPDE_B200_EXACT_B=1 switches the volume B.grad(q) terms from the reference's
// (defective, see kernels.cuh) evaluation to the intended matrix-vector product.
bool exact_b_product() {
  const char *e = getenv("PYPDE_B200_EXACT_B");
  return e && *e == '1';
}

bool fma_enabled() {
  const char *e = getenv("PYPDE_B200_FMA");
  return e && *e == '1';
}

// NVRTC: CUDA source -> LTO-IR
std::vector<char> nvrtc_to_ltoir(const std::string &src, const char *name,
                                 const std::vector<std::string> &defines) {
  const NvrtcApi &rt = nvrtc();
  nvrtcProgram prog;
  nvrtcResult r = rt.CreateProgram(&prog, src.c_str(), name, 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS)
    throw std::runtime_error(std::string("nvrtcCreateProgram: ") + rt.GetErrorString(r));
  std::vector<std::string> opts = {"--gpu-architecture=compute_100a", "-dlto", "-std=c++17",
                                   "-lineinfo", "--fmad=false"};
  if (fma_enabled())
    opts.back() = "--fmad=true";
  for (const std::string &d : defines)
    opts.push_back("-D" + d);
  std::vector<const char *> copts;
  for (const std::string &o : opts)
    copts.push_back(o.c_str());
  r = rt.CompileProgram(prog, (int)copts.size(), copts.data());
  size_t logn = 0;
  rt.GetProgramLogSize(prog, &logn);
  std::string log(logn, '\0');
  if (logn > 1)
    rt.GetProgramLog(prog, &log[0]);
  if (r != NVRTC_SUCCESS) {
    rt.DestroyProgram(&prog);
    throw std::runtime_error(std::string("pypde_b200: NVRTC failed on ") + name + ": " +
                             rt.GetErrorString(r) + "\n" + log);
  }
  size_t n = 0;
  rt.GetLTOIRSize(prog, &n);
  std::vector<char> out(n);
  rt.GetLTOIR(prog, out.data());
  rt.DestroyProgram(&prog);
  if (n == 0)
    throw std::runtime_error(std::string("pypde_b200: NVRTC produced no LTO-IR for ") + name);
  return out;
}

std::string jl_log(JitLinkHandle h) {
  const JitLinkApi &jl = jitlink();
  std::string s;
  size_t n = 0;
  if (jl.GetErrorLogSize(h, &n) == 0 && n > 1) {
    std::string e(n, '\0');
    jl.GetErrorLog(h, &e[0]);
    s += e.c_str();
  }
  n = 0;
  if (jl.GetInfoLogSize(h, &n) == 0 && n > 1) {
    std::string e(n, '\0');
    jl.GetInfoLog(h, &e[0]);
    s += e.c_str();
  }
  return s;
}

std::mutex g_cache_mutex;
std::map<std::string, std::vector<char>> g_cache;

// the opt-in wave-speed function (owned copy)
std::mutex g_ws_mutex;
std::string g_ws_image, g_ws_name;
int g_ws_kind = -1;

} // namespace

void set_wavespeed(const pypde_b200_devfn *L) {
  std::lock_guard<std::mutex> lk(g_ws_mutex);
  if (!L || !L->image || L->bytes == 0) {
    g_ws_image.clear();
    g_ws_name.clear();
    g_ws_kind = -1;
    return;
  }
  g_ws_image.assign((const char *)L->image, L->bytes);
  g_ws_name = L->name ? L->name : "user_L";
  g_ws_kind = L->kind;
}
bool wavespeed_set() {
  std::lock_guard<std::mutex> lk(g_ws_mutex);
  return g_ws_kind >= 0;
}
std::string wavespeed_key() {
  std::lock_guard<std::mutex> lk(g_ws_mutex);
  return g_ws_kind < 0 ? std::string() : std::to_string(g_ws_kind) + ":" + g_ws_image;
}

void choose_block_shapes(KernelConfig &c) {
  const int Nd = ipow(c.N, c.ndim);
  const int NT = c.N * Nd;
  const int NP = c.N * ipow(c.N, c.ndim - 1);
  if (NT > 1024)
    throw std::runtime_error("pypde_b200: order too high for this ndim (N^(ndim+1) > 1024)");
  // k_dg: NT threads per cell; (2+ndim)*NT*V doubles of shared memory per cell
  const size_t sm_cell = (size_t)(2 + c.ndim) * NT * c.V * 8;
  // one warp per block measured fastest (B200, 2-D Euler N=3: 4.3 ms vs 5.2 ms with
  // 9 cells / 243 threads): short blocks do not idle at the per-iteration barriers
  int cpb = 32 / NT;
  if (cpb < 1)
    cpb = 1;
  while (cpb > 1 && cpb * sm_cell > 64 * 1024)
    cpb--;
  if (cpb * sm_cell > 220 * 1024)
    throw std::runtime_error("pypde_b200: predictor working set exceeds shared memory "
                             "(N^(ndim+1) * V too large)");
  c.dg_cpb = cpb;
  // k_faces: NP threads per face; FLX_W*V doubles per thread
  const size_t sm_pt = (size_t)(c.useB ? 2 : 1) * c.V * 8;
  int fpb = 128 / NP;
  if (fpb < 1)
    fpb = 1;
  while (fpb > 1 && (size_t)fpb * NP * sm_pt > 48 * 1024)
    fpb--;
  c.faces_fpb = fpb;
  // k_dg_stiff: one warp per cell.  Shared memory per warp (kernels.cuh: NK_SMEM2) =
  // (3+ndim) n doubles of evaluation state + KS resident Krylov vectors of n doubles + the
  // packed triangular factor (32 columns) + Givens / least-squares data.  KS = 1: measured
  // on B200 (C3 at 512^2) 1 resident vector 28.4 ms per step, 2: 30.6, 3: 31.8, 8: 36.3,
  // all 31: 61.9 — the inner solves of these cells run ~26 steps deep, the kernel lives on
  // occupancy (16 warps per SM at 128 registers) and on L1 for the rest of the basis, and
  // every resident vector takes from both.
  {
    const size_t n = (size_t)NT * c.V;
    if (const char *e = getenv("PYPDE_B200_STIFF_STATS"))
      c.stiff_stats = *e == '1';
    int ks = 1;
    if (const char *e = getenv("PYPDE_B200_STIFF_KS"))
      ks = atoi(e) < 1 ? 1 : (atoi(e) > 40 ? 40 : atoi(e));
    c.stiff_ks = ks;
    auto sm_warp = [&]() {
      return ((size_t)(3 + c.ndim + c.stiff_ks) * n + (size_t)32 * 35 / 2 + 5 * 41 + 2) * 8;
    };
    int wpb = 4;
    if (const char *e = getenv("PYPDE_B200_STIFF_WPB"))
      wpb = atoi(e) < 1 ? 1 : atoi(e);
    while (c.stiff_ks > 1 && sm_warp() > 220 * 1024)
      c.stiff_ks--;
    while (wpb > 1 && wpb * sm_warp() > 110 * 1024)
      wpb--;
    c.stiff_wpb = wpb;
    // 4 blocks of 4 warps per SM <-> 128 registers per thread: no spills for V <= 8 (reactive
    // Euler: 158 uncapped); larger systems (GPR, V = 17: 254 uncapped) keep the full budget
    c.stiff_minblocks = (c.V <= 8 && wpb == 4) ? 4 : 1;
    if (const char *e = getenv("PYPDE_B200_STIFF_MINBLOCKS"))
      c.stiff_minblocks = atoi(e);
  }
  if (const char *e = getenv("PYPDE_B200_W3_TILE"))
    sscanf(e, "%d,%d,%d", &c.w3_ti, &c.w3_tj, &c.w3_tk);
  // V > 5 (the eigen-solves work in local memory, one long dependency chain per thread): 20
  // warps per SM at 96 registers measured 2-6 % ahead of 16 at 128 (profiles/r2_occupancy_sweep.txt)
  if (c.V > 5)
    c.ws_block = 640;
  // tuning overrides (experiments)
  if (const char *e = getenv("PYPDE_B200_WS_BLOCK"))
    c.ws_block = atoi(e);
  if (const char *e = getenv("PYPDE_B200_WS_MINBLOCKS"))
    c.ws_minblocks = atoi(e);
  if (const char *e = getenv("PYPDE_B200_FF_BLOCK"))
    c.ff_block = atoi(e);
  if (const char *e = getenv("PYPDE_B200_FF_MINBLOCKS"))
    c.ff_minblocks = atoi(e);
  // k_faces_side with a second-order flux: one 512-thread block per SM measured best
  // (15.8 ms against 17.1 for 256 x 2 and 17.9 for 128 x 4 at C5 32 x 128^2)
  if (c.secondOrder) {
    c.fs_block = 512;
    c.fs_minblocks = 1;
  }
  if (const char *e = getenv("PYPDE_B200_FS_BLOCK"))
    c.fs_block = atoi(e);
  if (const char *e = getenv("PYPDE_B200_FS_MINBLOCKS"))
    c.fs_minblocks = atoi(e);
  if (const char *e = getenv("PYPDE_B200_DG_CPB"))
    c.dg_cpb = atoi(e);
  if (const char *e = getenv("PYPDE_B200_FACES_FPB"))
    c.faces_fpb = atoi(e);
}

std::vector<std::string> specialisation_defines(const KernelConfig &c) {
  auto kv = [](const char *k, int v) {
    char b[64];
    snprintf(b, sizeof b, "%s=%d", k, v);
    return std::string(b);
  };
  std::vector<std::string> defs = {kv("PDE_NDIM", c.ndim),
          kv("PDE_N", c.N),
          kv("PDE_V", c.V),
          kv("PDE_FLUX", c.flux),
          kv("PDE_STIFF", c.stiff ? 1 : 0),
          kv("PDE_USE_F", c.useF ? 1 : 0),
          kv("PDE_USE_B", c.useB ? 1 : 0),
          kv("PDE_USE_S", c.useS ? 1 : 0),
          kv("PDE_SECOND_ORDER", c.secondOrder ? 1 : 0),
          kv("PDE_USE_L", c.useL ? 1 : 0),
          kv("PDE_EXACT_B_PRODUCT", exact_b_product() ? 1 : 0),
          kv("PDE_EIG_QR_ONLY", getenv("PYPDE_B200_EIG_QR_ONLY") ? 1 : 0),
          kv("PDE_DG_CPB", c.dg_cpb),
          kv("PDE_FACES_FPB", c.faces_fpb),
          kv("PDE_STIFF_WPB", c.stiff_wpb),
          kv("PDE_W3_TI", c.w3_ti),
          kv("PDE_W3_TJ", c.w3_tj),
          kv("PDE_W3_TK", c.w3_tk),
          kv("PDE_STIFF_KS", c.stiff_ks),
          kv("PDE_STIFF_MINBLOCKS", c.stiff_minblocks),
          kv("PDE_STIFF_STATS", c.stiff_stats ? 1 : 0),
          kv("PDE_WS_BLOCK", c.ws_block),
          kv("PDE_WS_MINBLOCKS", c.ws_minblocks),
          kv("PDE_FF_BLOCK", c.ff_block),
          kv("PDE_FF_MINBLOCKS", c.ff_minblocks),
          kv("PDE_FS_BLOCK", c.fs_block),
          kv("PDE_FS_MINBLOCKS", c.fs_minblocks)};
  // tuning experiments: PYPDE_B200_EXTRA_DEFINES="PDE_X=1;PDE_Y=0"
  if (const char *e = getenv("PYPDE_B200_EXTRA_DEFINES")) {
    std::string all(e);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t q = all.find(';', pos);
      if (q == std::string::npos)
        q = all.size();
      if (q > pos)
        defs.push_back(all.substr(pos, q - pos));
      pos = q + 1;
    }
  }
  return defs;
}

std::string specialised_source(const KernelConfig &c) {
  std::string s = tables_cuda_source(make_tables(c.N));
  s += KERNEL_SOURCE;
  return s;
}

std::vector<char> build_cubin(const KernelConfig &cfg, const pypde_b200_devfn *F,
                              const pypde_b200_devfn *B, const pypde_b200_devfn *S) {
  const std::string src = specialised_source(cfg);
  const std::vector<std::string> defs = specialisation_defines(cfg);
  const pypde_b200_devfn *fn[3] = {cfg.useF ? F : nullptr, cfg.useB ? B : nullptr,
                                   cfg.useS ? S : nullptr};
  const char *fn_names[3] = {"user_F", "user_B", "user_S"};
  // the opt-in wave-speed function travels as a fourth image
  std::string ws_image, ws_name;
  int ws_kind = -1;
  if (cfg.useL) {
    std::lock_guard<std::mutex> lk(g_ws_mutex);
    ws_image = g_ws_image;
    ws_name = g_ws_name;
    ws_kind = g_ws_kind;
    if (ws_kind < 0)
      throw std::runtime_error("pypde_b200: configuration wants user_L but no wave-speed function is set");
  }

  Hasher h;
  h.add(src);
  for (const std::string &d : defs)
    h.add(d);
  h.add(fma_enabled() ? "fma1" : "fma0");
  {
    // a toolkit upgrade must not reuse cubins of the old compiler / linker
    int v[4] = {0, 0, 0, 0};
    unsigned lv[2] = {0, 0};
    nvrtc().Version(&v[0], &v[1]);
    jitlink().Version(&lv[0], &lv[1]);
    v[2] = (int)lv[0];
    v[3] = (int)lv[1];
    h.add(v, sizeof v);
  }
  for (int i = 0; i < 3; i++) {
    if (!fn[i]) {
      h.add("-");
      continue;
    }
    if (!fn[i]->image || fn[i]->bytes == 0)
      throw std::runtime_error(std::string("pypde_b200: empty device-function image for ") +
                               fn_names[i]);
    h.add(&fn[i]->kind, sizeof(int));
    h.add(fn[i]->image, fn[i]->bytes);
  }
  if (cfg.useL) {
    h.add(&ws_kind, sizeof(int));
    h.add(ws_image);
  }
  const std::string key = h.hex();
  {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    auto it = g_cache.find(key);
    if (it != g_cache.end())
      return it->second;
  }
  const std::string dir = getenv("PYPDE_B200_NO_DISK_CACHE") ? std::string() : cache_dir();
  const bool use_disk = !dir.empty();
  const std::string path = dir + "/" + key + ".cubin";
  std::vector<char> cubin;
  if (use_disk && read_file(path, cubin)) {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    g_cache[key] = cubin;
    return cubin;
  }
  if (!getenv("PYPDE_B200_NO_DISK_CACHE")) {
    const std::string shipped = shipped_cache_dir();
    if (getenv("PYPDE_B200_CACHE_DEBUG"))
      fprintf(stderr, "pypde_b200: cubin %s: not in '%s'; shipped cache '%s'\n", key.c_str(),
              dir.c_str(), shipped.c_str());
    if (!shipped.empty() && read_file(shipped + "/" + key + ".cubin", cubin)) {
      std::lock_guard<std::mutex> lk(g_cache_mutex);
      g_cache[key] = cubin;
      return cubin;
    }
  }

  // 1. our kernels -> LTO-IR
  std::vector<char> kernels_ir = nvrtc_to_ltoir(src, "pypde_b200_kernels.cu", defs);

  // 2. link with the user functions
  const JitLinkApi &jl = jitlink();
  std::vector<const char *> lopts = {"-arch=sm_100a", "-lto", "-O3", "-lineinfo"};
  lopts.push_back(fma_enabled() ? "-fma=1" : "-fma=0");
  JitLinkHandle lh = nullptr;
  int rc = jl.Create(&lh, (unsigned)lopts.size(), lopts.data());
  if (rc != 0)
    throw std::runtime_error("pypde_b200: nvJitLinkCreate failed (code " + std::to_string(rc) +
                             ")" + (lh ? "\n" + jl_log(lh) : ""));
  auto fail = [&](const std::string &what, int code) {
    std::string log = jl_log(lh);
    jl.Destroy(&lh);
    throw std::runtime_error("pypde_b200: " + what + " failed (code " + std::to_string(code) +
                             ")\n" + log);
  };
  rc = jl.AddData(lh, JITLINK_INPUT_LTOIR, kernels_ir.data(), kernels_ir.size(),
                  "pypde_b200_kernels");
  if (rc != 0)
    fail("nvJitLinkAddData(kernels)", rc);
  for (int i = 0; i < 3; i++) {
    if (!fn[i])
      continue;
    const char *label = fn[i]->name ? fn[i]->name : fn_names[i];
    if (fn[i]->kind == PYPDE_B200_LTOIR) {
      rc = jl.AddData(lh, JITLINK_INPUT_LTOIR, fn[i]->image, fn[i]->bytes, label);
    } else if (fn[i]->kind == PYPDE_B200_PTX) {
      rc = jl.AddData(lh, JITLINK_INPUT_PTX, fn[i]->image, fn[i]->bytes, label);
    } else if (fn[i]->kind == PYPDE_B200_CUDA_SOURCE) {
      std::string usrc((const char *)fn[i]->image,
                       strnlen((const char *)fn[i]->image, fn[i]->bytes));
      std::vector<char> ir;
      try {
        ir = nvrtc_to_ltoir(usrc, label, defs);
      } catch (...) {
        jl.Destroy(&lh);
        throw;
      }
      rc = jl.AddData(lh, JITLINK_INPUT_LTOIR, ir.data(), ir.size(), label);
    } else {
      jl.Destroy(&lh);
      throw std::runtime_error(std::string("pypde_b200: unknown device-function kind for ") +
                               fn_names[i]);
    }
    if (rc != 0)
      fail(std::string("nvJitLinkAddData(") + fn_names[i] + ")", rc);
  }
  if (cfg.useL) {
    if (ws_kind == PYPDE_B200_LTOIR) {
      rc = jl.AddData(lh, JITLINK_INPUT_LTOIR, ws_image.data(), ws_image.size(), ws_name.c_str());
    } else if (ws_kind == PYPDE_B200_PTX) {
      rc = jl.AddData(lh, JITLINK_INPUT_PTX, ws_image.data(), ws_image.size(), ws_name.c_str());
    } else {
      std::vector<char> ir;
      try {
        ir = nvrtc_to_ltoir(std::string(ws_image.c_str()), ws_name.c_str(), defs);
      } catch (...) {
        jl.Destroy(&lh);
        throw;
      }
      rc = jl.AddData(lh, JITLINK_INPUT_LTOIR, ir.data(), ir.size(), ws_name.c_str());
    }
    if (rc != 0)
      fail("nvJitLinkAddData(user_L)", rc);
  }
  rc = jl.Complete(lh);
  if (rc != 0)
    fail("nvJitLinkComplete (is a user_F/user_B/user_S/user_L symbol missing or mis-typed?)", rc);
  size_t n = 0;
  rc = jl.GetLinkedCubinSize(lh, &n);
  if (rc != 0 || n == 0)
    fail("nvJitLinkGetLinkedCubinSize", rc);
  cubin.resize(n);
  rc = jl.GetLinkedCubin(lh, cubin.data());
  if (rc != 0)
    fail("nvJitLinkGetLinkedCubin", rc);
  jl.Destroy(&lh);

  if (use_disk)
    write_file_atomic(path, cubin);
  std::lock_guard<std::mutex> lk(g_cache_mutex);
  g_cache[key] = cubin;
  return cubin;
}

} // namespace pypde
