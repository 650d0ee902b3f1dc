// Lazily dlopen()ed toolchain and driver libraries.
//
// libpypde.so must load on a machine without a GPU driver (CPU test-suite,
// build check), and inside a Python process that may already hold *other*
// copies of libnvrtc / libnvJitLink (torch's pip wheels ship 12.8 ones whose
// nvJitLink rejects 12.9 LTO-IR).  So nothing here is a DT_NEEDED dependency:
// NVRTC and nvJitLink are opened by absolute path from the CUDA toolkit, the
// driver (libcuda.so.1) and NCCL by soname, all on first use.
#pragma once
#include <cuda.h>
#include <nvrtc.h>
#include <stddef.h>
#include <string>

namespace pypde {

struct DriverApi {
  CUresult (*Init)(unsigned);
  CUresult (*DeviceGet)(CUdevice *, int);
  CUresult (*DeviceGetCount)(int *);
  CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
  CUresult (*DevicePrimaryCtxRetain)(CUcontext *, CUdevice);
  CUresult (*CtxGetCurrent)(CUcontext *);
  CUresult (*CtxSetCurrent)(CUcontext);
  CUresult (*CtxGetDevice)(CUdevice *);
  CUresult (*MemAlloc)(CUdeviceptr *, size_t);
  CUresult (*MemFree)(CUdeviceptr);
  CUresult (*MemAllocHost)(void **, size_t);
  CUresult (*MemFreeHost)(void *);
  CUresult (*MemcpyHtoDAsync)(CUdeviceptr, const void *, size_t, CUstream);
  CUresult (*MemcpyDtoHAsync)(void *, CUdeviceptr, size_t, CUstream);
  CUresult (*MemcpyDtoDAsync)(CUdeviceptr, CUdeviceptr, size_t, CUstream);
  CUresult (*MemsetD8Async)(CUdeviceptr, unsigned char, size_t, CUstream);
  CUresult (*StreamCreate)(CUstream *, unsigned);
  CUresult (*StreamDestroy)(CUstream);
  CUresult (*StreamSynchronize)(CUstream);
  CUresult (*ModuleLoadData)(CUmodule *, const void *);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*FuncGetAttribute)(int *, CUfunction_attribute, CUfunction);
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, CUfunction, int, size_t);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, unsigned, CUstream, void **, void **);
  CUresult (*GetErrorString)(CUresult, const char **);
  CUresult (*EventCreate)(CUevent *, unsigned);
  CUresult (*EventDestroy)(CUevent);
  CUresult (*EventRecord)(CUevent, CUstream);
  CUresult (*EventSynchronize)(CUevent);
  CUresult (*EventElapsedTime)(float *, CUevent, CUevent);
  CUresult (*StreamWaitEvent)(CUstream, CUevent, unsigned);
  CUresult (*StreamBeginCapture)(CUstream, CUstreamCaptureMode);
  CUresult (*StreamEndCapture)(CUstream, CUgraph *);
  CUresult (*GraphInstantiate)(CUgraphExec *, CUgraph, unsigned long long);
  CUresult (*GraphLaunch)(CUgraphExec, CUstream);
  CUresult (*GraphExecDestroy)(CUgraphExec);
  CUresult (*GraphDestroy)(CUgraph);
  CUresult (*TensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                   const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
};

struct NvrtcApi {
  nvrtcResult (*Version)(int *, int *);
  nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int,
                               const char *const *, const char *const *);
  nvrtcResult (*DestroyProgram)(nvrtcProgram *);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
  nvrtcResult (*GetLTOIRSize)(nvrtcProgram, size_t *);
  nvrtcResult (*GetLTOIR)(nvrtcProgram, char *);
  const char *(*GetErrorString)(nvrtcResult);
};

// nvJitLink.h maps its entry points to versioned names; the library also
// exports the plain ones, which is what we bind.
typedef struct nvJitLink_opaque *JitLinkHandle;
enum { JITLINK_INPUT_CUBIN = 1, JITLINK_INPUT_PTX = 2, JITLINK_INPUT_LTOIR = 3 };
struct JitLinkApi {
  int (*Version)(unsigned *, unsigned *);
  int (*Create)(JitLinkHandle *, unsigned, const char **);
  int (*Destroy)(JitLinkHandle *);
  int (*AddData)(JitLinkHandle, int, const void *, size_t, const char *);
  int (*Complete)(JitLinkHandle);
  int (*GetLinkedCubinSize)(JitLinkHandle, size_t *);
  int (*GetLinkedCubin)(JitLinkHandle, void *);
  int (*GetErrorLogSize)(JitLinkHandle, size_t *);
  int (*GetErrorLog)(JitLinkHandle, char *);
  int (*GetInfoLogSize)(JitLinkHandle, size_t *);
  int (*GetInfoLog)(JitLinkHandle, char *);
};

// NCCL: only the handful of calls the slab exchange needs.
typedef struct ncclComm *NcclComm;
struct NcclUniqueId {
  char internal[128];
};
struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *);
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*Send)(const void *, size_t, int /*dtype*/, int /*peer*/, NcclComm, CUstream);
  int (*Recv)(void *, size_t, int, int, NcclComm, CUstream);
  int (*AllReduce)(const void *, void *, size_t, int /*dtype*/, int /*op*/, NcclComm, CUstream);
  const char *(*GetErrorString)(int);
};
enum { NCCL_FLOAT64 = 8, NCCL_UINT64 = 5, NCCL_INT32 = 2, NCCL_MAX = 2 };

// Each throws std::runtime_error with a clear message if the library or a
// symbol is missing.
const DriverApi &driver();
const NvrtcApi &nvrtc();
const JitLinkApi &jitlink();
const NcclApi &nccl();

std::string cuda_home();

} // namespace pypde
