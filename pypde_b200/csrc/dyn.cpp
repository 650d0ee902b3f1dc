#include "dyn.h"

#include <dlfcn.h>
#include <mutex>
#include <stdexcept>
#include <stdlib.h>
#include <vector>

namespace pypde {

std::string cuda_home() {
  const char *e = getenv("PYPDE_B200_CUDA_HOME");
  if (e && *e)
    return e;
  e = getenv("CUDA_HOME");
  if (e && *e)
    return e;
  return "/usr/local/cuda";
}

namespace {

void *open_first(const std::vector<std::string> &names, const char *what) {
  std::string tried;
  for (const std::string &n : names) {
    void *h = dlopen(n.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (h)
      return h;
    tried += "\n  " + n + ": " + dlerror();
  }
  throw std::runtime_error(std::string("pypde_b200: cannot load ") + what + "; tried:" + tried);
}

template <typename T> void bind(void *h, T &fn, const char *sym, const char *lib) {
  void *p = dlsym(h, sym);
  if (!p)
    throw std::runtime_error(std::string("pypde_b200: symbol ") + sym + " missing in " + lib);
  fn = reinterpret_cast<T>(p);
}

} // namespace

const DriverApi &driver() {
  static DriverApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    try {
      void *h = open_first({"libcuda.so.1", "libcuda.so"}, "the CUDA driver (libcuda.so.1) — "
                                                           "no GPU driver on this machine?");
#define B(field, sym) bind(h, api.field, sym, "libcuda")
      B(Init, "cuInit");
      B(DeviceGet, "cuDeviceGet");
      B(DeviceGetCount, "cuDeviceGetCount");
      B(DeviceGetAttribute, "cuDeviceGetAttribute");
      B(DevicePrimaryCtxRetain, "cuDevicePrimaryCtxRetain");
      B(CtxGetCurrent, "cuCtxGetCurrent");
      B(CtxSetCurrent, "cuCtxSetCurrent");
      B(CtxGetDevice, "cuCtxGetDevice");
      B(MemAlloc, "cuMemAlloc_v2");
      B(MemFree, "cuMemFree_v2");
      B(MemAllocHost, "cuMemAllocHost_v2");
      B(MemFreeHost, "cuMemFreeHost");
      B(MemcpyHtoDAsync, "cuMemcpyHtoDAsync_v2");
      B(MemcpyDtoHAsync, "cuMemcpyDtoHAsync_v2");
      B(MemcpyDtoDAsync, "cuMemcpyDtoDAsync_v2");
      B(MemsetD8Async, "cuMemsetD8Async");
      B(StreamCreate, "cuStreamCreate");
      B(StreamDestroy, "cuStreamDestroy_v2");
      B(StreamSynchronize, "cuStreamSynchronize");
      B(ModuleLoadData, "cuModuleLoadData");
      B(ModuleUnload, "cuModuleUnload");
      B(ModuleGetFunction, "cuModuleGetFunction");
      B(FuncSetAttribute, "cuFuncSetAttribute");
      B(FuncGetAttribute, "cuFuncGetAttribute");
      B(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
      B(LaunchKernel, "cuLaunchKernel");
      B(GetErrorString, "cuGetErrorString");
      B(EventCreate, "cuEventCreate");
      B(EventDestroy, "cuEventDestroy_v2");
      B(EventRecord, "cuEventRecord");
      B(EventSynchronize, "cuEventSynchronize");
      B(EventElapsedTime, "cuEventElapsedTime");
      B(StreamWaitEvent, "cuStreamWaitEvent");
      B(TensorMapEncodeTiled, "cuTensorMapEncodeTiled");
      B(StreamBeginCapture, "cuStreamBeginCapture_v2");
      B(StreamEndCapture, "cuStreamEndCapture");
      B(GraphInstantiate, "cuGraphInstantiateWithFlags");
      B(GraphLaunch, "cuGraphLaunch");
      B(GraphExecDestroy, "cuGraphExecDestroy");
      B(GraphDestroy, "cuGraphDestroy");
#undef B
      CUresult r = api.Init(0);
      if (r != CUDA_SUCCESS) {
        const char *s = nullptr;
        api.GetErrorString(r, &s);
        throw std::runtime_error(std::string("pypde_b200: cuInit failed: ") + (s ? s : "?"));
      }
    } catch (const std::exception &e) {
      err = e.what();
    }
  });
  if (!err.empty())
    throw std::runtime_error(err);
  return api;
}

const NvrtcApi &nvrtc() {
  static NvrtcApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    try {
      std::string home = cuda_home();
      void *h = open_first({home + "/lib64/libnvrtc.so.12", home + "/lib64/libnvrtc.so",
                            "libnvrtc.so.12"},
                           "NVRTC");
#define B(field, sym) bind(h, api.field, sym, "libnvrtc")
      B(Version, "nvrtcVersion");
      B(CreateProgram, "nvrtcCreateProgram");
      B(DestroyProgram, "nvrtcDestroyProgram");
      B(CompileProgram, "nvrtcCompileProgram");
      B(GetProgramLogSize, "nvrtcGetProgramLogSize");
      B(GetProgramLog, "nvrtcGetProgramLog");
      B(GetLTOIRSize, "nvrtcGetLTOIRSize");
      B(GetLTOIR, "nvrtcGetLTOIR");
      B(GetErrorString, "nvrtcGetErrorString");
#undef B
    } catch (const std::exception &e) {
      err = e.what();
    }
  });
  if (!err.empty())
    throw std::runtime_error(err);
  return api;
}

const JitLinkApi &jitlink() {
  static JitLinkApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    try {
      std::string home = cuda_home();
      void *h = open_first({home + "/lib64/libnvJitLink.so.12", home + "/lib64/libnvJitLink.so",
                            "libnvJitLink.so.12"},
                           "nvJitLink");
#define B(field, sym) bind(h, api.field, sym, "libnvJitLink")
      B(Version, "nvJitLinkVersion");
      B(Create, "nvJitLinkCreate");
      B(Destroy, "nvJitLinkDestroy");
      B(AddData, "nvJitLinkAddData");
      B(Complete, "nvJitLinkComplete");
      B(GetLinkedCubinSize, "nvJitLinkGetLinkedCubinSize");
      B(GetLinkedCubin, "nvJitLinkGetLinkedCubin");
      B(GetErrorLogSize, "nvJitLinkGetErrorLogSize");
      B(GetErrorLog, "nvJitLinkGetErrorLog");
      B(GetInfoLogSize, "nvJitLinkGetInfoLogSize");
      B(GetInfoLog, "nvJitLinkGetInfoLog");
#undef B
    } catch (const std::exception &e) {
      err = e.what();
    }
  });
  if (!err.empty())
    throw std::runtime_error(err);
  return api;
}

const NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    try {
      // inside a torch process this resolves to the already loaded bundled NCCL
      void *h = open_first({"libnccl.so.2", "libnccl.so"}, "NCCL");
#define B(field, sym) bind(h, api.field, sym, "libnccl")
      B(GetUniqueId, "ncclGetUniqueId");
      B(CommInitRank, "ncclCommInitRank");
      B(CommDestroy, "ncclCommDestroy");
      B(GroupStart, "ncclGroupStart");
      B(GroupEnd, "ncclGroupEnd");
      B(Send, "ncclSend");
      B(Recv, "ncclRecv");
      B(AllReduce, "ncclAllReduce");
      B(GetErrorString, "ncclGetErrorString");
#undef B
    } catch (const std::exception &e) {
      err = e.what();
    }
  });
  if (!err.empty())
    throw std::runtime_error(err);
  return api;
}

} // namespace pypde
