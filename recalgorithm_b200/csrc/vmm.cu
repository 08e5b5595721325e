// Peer-mappable device memory from the CUDA virtual-memory-management API (cuMemCreate / cuMemMap), shared between the
// one-process-per-GPU ranks as POSIX file descriptors.  Used for the table shard of the row-sharded path: the
// random 128-byte peer reads of the gather are sensitive to HOW the peer mapping was made -- a legacy CUDA-IPC mapping of a
// 32 GB shard ran at 7 GB/s (round 1), and with VMM mappings the page granularity / alignment of the allocation decides the
// reach of the peer TLB.  This file lets the host side choose both explicitly.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "ctr_common.cuh"

namespace ctr {
namespace {

struct Driver {
  CUresult (*getGran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*exportH)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*importH)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  bool ok = false;
};

template <typename F>
bool load(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
    return false;
  fn = reinterpret_cast<F>(p);
  return true;
}

const Driver& driver() {
  static Driver d = [] {
    Driver x;
    x.ok = load(x.getGran, "cuMemGetAllocationGranularity") && load(x.create, "cuMemCreate") && load(x.release, "cuMemRelease") &&
           load(x.reserve, "cuMemAddressReserve") && load(x.addrFree, "cuMemAddressFree") && load(x.map, "cuMemMap") &&
           load(x.unmap, "cuMemUnmap") && load(x.setAccess, "cuMemSetAccess") && load(x.exportH, "cuMemExportToShareableHandle") &&
           load(x.importH, "cuMemImportFromShareableHandle");
    return x;
  }();
  return d;
}

struct Mapping { CUmemGenericAllocationHandle h; size_t bytes; };
std::mutex g_mu;
std::unordered_map<unsigned long long, Mapping> g_maps;

CUmemAllocationProp prop_for(int dev) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

#define CTR_CU(call)                                                   \
  do {                                                                 \
    CUresult r__ = (call);                                             \
    if (r__ != CUDA_SUCCESS) {                                         \
      set_error("%s failed with CUresult %d", #call, (int)r__);        \
      return CTR_ERR_CUDA;                                             \
    }                                                                  \
  } while (0)

int map_handle(const Driver& D, CUmemGenericAllocationHandle h, size_t bytes, size_t align, int dev, void** ptr) {
  CUdeviceptr va = 0;
  CTR_CU(D.reserve(&va, bytes, align, 0, 0));
  CTR_CU(D.map(va, bytes, 0, h, 0));
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CTR_CU(D.setAccess(va, bytes, &acc, 1));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_maps[(unsigned long long)va] = Mapping{h, bytes};
  }
  *ptr = reinterpret_cast<void*>(va);
  return CTR_OK;
}

}  // namespace
}  // namespace ctr

using namespace ctr;

extern "C" int ctr_vmm_granularity(int64_t* minimum, int64_t* recommended) {
  const Driver& D = driver();
  CTR_REQUIRE(D.ok, "ctr_vmm_granularity: the CUDA driver lacks the virtual-memory-management entry points");
  int dev = 0;
  CTR_CUDA(cudaGetDevice(&dev));
  CTR_CUDA(cudaFree(nullptr));                                  // make sure the primary context exists
  const CUmemAllocationProp p = prop_for(dev);
  size_t g = 0;
  if (minimum) { CTR_CU(D.getGran(&g, &p, CU_MEM_ALLOC_GRANULARITY_MINIMUM)); *minimum = (int64_t)g; }
  if (recommended) { CTR_CU(D.getGran(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED)); *recommended = (int64_t)g; }
  return CTR_OK;
}

extern "C" int ctr_vmm_alloc(int64_t bytes, int64_t align, void** ptr, int* fd, int64_t* mapped_bytes) {
  const Driver& D = driver();
  CTR_REQUIRE(D.ok, "ctr_vmm_alloc: the CUDA driver lacks the virtual-memory-management entry points");
  CTR_REQUIRE(bytes > 0 && ptr && fd && mapped_bytes, "ctr_vmm_alloc: bad arguments");
  int dev = 0;
  CTR_CUDA(cudaGetDevice(&dev));
  CTR_CUDA(cudaFree(nullptr));
  const CUmemAllocationProp p = prop_for(dev);
  size_t gran = 0;
  CTR_CU(D.getGran(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  size_t a = align > 0 ? (size_t)align : gran;
  CTR_REQUIRE(a % gran == 0, "ctr_vmm_alloc: align=%lld must be a multiple of the granularity %zu", (long long)align, gran);
  const size_t size = ((size_t)bytes + a - 1) / a * a;         // size AND address aligned to `a`: lets the driver use its largest pages
  CUmemGenericAllocationHandle h;
  CTR_CU(D.create(&h, size, &p, 0));
  int rc = map_handle(D, h, size, a, dev, ptr);
  if (rc) return rc;
  int out_fd = -1;
  CTR_CU(D.exportH(&out_fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd = out_fd;
  *mapped_bytes = (int64_t)size;
  return CTR_OK;
}

extern "C" int ctr_vmm_import(int fd, int64_t mapped_bytes, int64_t align, void** ptr) {
  const Driver& D = driver();
  CTR_REQUIRE(D.ok, "ctr_vmm_import: the CUDA driver lacks the virtual-memory-management entry points");
  CTR_REQUIRE(fd >= 0 && mapped_bytes > 0 && ptr, "ctr_vmm_import: bad arguments");
  int dev = 0;
  CTR_CUDA(cudaGetDevice(&dev));
  CTR_CUDA(cudaFree(nullptr));
  CUmemGenericAllocationHandle h;
  CTR_CU(D.importH(&h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  size_t gran = 0;
  const CUmemAllocationProp p = prop_for(dev);
  CTR_CU(D.getGran(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  return map_handle(D, h, (size_t)mapped_bytes, align > 0 ? (size_t)align : gran, dev, ptr);
}

extern "C" int ctr_vmm_free(void* ptr) {
  const Driver& D = driver();
  if (ptr == nullptr) return CTR_OK;
  CTR_REQUIRE(D.ok, "ctr_vmm_free: the CUDA driver lacks the virtual-memory-management entry points");
  Mapping m;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find((unsigned long long)ptr);
    CTR_REQUIRE(it != g_maps.end(), "ctr_vmm_free: %p was not mapped by ctr_vmm_alloc / ctr_vmm_import", ptr);
    m = it->second;
    g_maps.erase(it);
  }
  CTR_CU(D.unmap((CUdeviceptr)ptr, m.bytes));
  CTR_CU(D.addrFree((CUdeviceptr)ptr, m.bytes));
  CTR_CU(D.release(m.h));
  return CTR_OK;
}
