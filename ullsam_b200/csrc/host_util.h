// ullsam_b200 — host-side helpers shared by the kernel launchers.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <utility>

namespace ullsam {

// Kernel launch, optionally with programmatic dependent launch (PDL): the ~2-3 us between two dependent kernels of a
// stream (launch latency + ramp-up) overlap the tail of the predecessor; the kernel must call pdl_wait() (common.cuh)
// before it touches global memory.  OFF by default (plain stream order), ULLSAM_PDL=1 turns the attribute on: measured on
// B200 it does not move the 36 ms step (36.21 ms with, 36.10 ms without, same box — the kernels are ~300 us long and
// persistent, there is no launch gap worth hiding), and one bench run with it on ended in the mbarrier watchdog of the
// global attention kernel (never seen in plain stream order; not understood).  cluster_x > 1 adds a
// cluster dimension.
inline bool pdl_enabled() {
  static const bool on = getenv("ULLSAM_PDL") != nullptr && atoi(getenv("ULLSAM_PDL")) != 0;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// "Has this kernel been configured on the CURRENT device yet?"  cudaFuncSetAttribute is per device (context), so a
// process that drives several GPUs has to set MaxDynamicSharedMemorySize on each of them; one flag bit per device,
// updated atomically (launchers may be called from several host threads, one per handle).
struct PerDeviceOnce {
  std::atomic<uint64_t> done{0};
  // returns true exactly when the caller still has to configure the current device (and then must call mark())
  bool needed(int* dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev;
    return dev < 0 || dev >= 64 || !(done.load(std::memory_order_acquire) & (1ull << dev));
  }
  void mark(int dev) {
    if (dev >= 0 && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
  }
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per device; returns an error string or nullptr.
template <typename Kern>
inline const char* set_max_smem_once(PerDeviceOnce& once, Kern kern, int bytes, int carveout_percent = -1) {
  int dev = 0;
  if (!once.needed(&dev)) return nullptr;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  if (carveout_percent >= 0) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_percent);
    if (e != cudaSuccess) return cudaGetErrorString(e);
  }
  once.mark(dev);
  return nullptr;
}

// RAII: make `device` current for the duration of an ABI call and restore the caller's device afterwards
// (PyTorch tracks the current device per thread; an ABI call must not change it behind its back).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t status = cudaSuccess;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != device) {
      status = cudaSetDevice(device);
      switched = (status == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace ullsam
