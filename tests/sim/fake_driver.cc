// fake_driver.cc -- the sixteen CUDA driver entry points csrc/vmm_arena.cc asks cudaGetDriverEntryPoint for, on the
// simulated runtime (test infrastructure): virtual-memory allocations are memfds, "exporting" one is dup(), importing
// is fstat() (the inode names the allocation on every process), mapping is mmap(MAP_FIXED) into a PROT_NONE
// reservation -- so the engine's VMM arena, its descriptor exchange between processes (pidfd_getfd or SCM_RIGHTS:
// the real system calls) and its collective fallback logic run on a CPU as they are.
//
// A multicast object is a 4 KiB memfd that records which allocation every device bound.  Its mapping stays PROT_NONE
// -- nothing but multimem instructions may touch a multicast address -- and the kernels run from source
// (tests/sim/host_emu.h) resolve such an address into the n local mappings of the bound allocations:
// multimem.ld_reduce adds them (in device order; the switch's order is unspecified), multimem.st stores to all.
// MXKV_SIM_NO_VMM=1: report the entry points missing (the arena then falls back to cudaMalloc + cudaIpc as before).
#include <cuda.h>
#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace sim {

namespace {

constexpr size_t kGranularity = size_t(2) << 20;
constexpr size_t kCtrlBytes = 4096;
constexpr uint32_t kCtrlMagic = 0x6d636173;     // a multicast object's control page starts with this
constexpr int kMaxDev = 8;

struct McCtrl {                                  // shared between the processes through the object's memfd
  uint32_t magic, num_devices;
  uint64_t bytes;
  uint32_t added;                                // bit d: device d has joined
  struct Slot { uint64_t ino, dev; uint64_t mem_offset, mc_offset, size; uint32_t bound; uint32_t pad_; } slot[kMaxDev];
};

struct Handle {
  bool multicast = false;
  int fd = -1;
  size_t bytes = 0;
  int device = -1;                               // allocations created here: the device they live on
  uint64_t ino = 0, dev = 0;
  McCtrl* ctrl = nullptr;
};
struct Mapping {
  uintptr_t va = 0;
  size_t bytes = 0;
  bool multicast = false;
  uint64_t ino = 0, dev = 0;
  McCtrl* ctrl = nullptr;
};

std::mutex g_mu;
std::map<uint64_t, Handle> g_handles;
uint64_t g_next = 1;
std::vector<Mapping> g_maps;
std::map<uintptr_t, size_t> g_reserved;

// MXKV_SIM_VMM_FAIL="<entry point>[#k]:<rank>": that entry point fails on that rank (its k-th call; every call without
// #k) -- what the arena's agree-or-fall-back protocol exists for (tests/test_sim_host_logic.py)
bool Inject(const char* name) {
  static const char* spec = getenv("MXKV_SIM_VMM_FAIL");
  if (spec == nullptr) return false;
  static std::map<std::string, int> calls;
  std::string s = spec;
  const size_t colon = s.rfind(':');
  if (colon == std::string::npos) return false;
  const char* rank_env = getenv("RANK");
  if (rank_env == nullptr || atoi(rank_env) != atoi(s.c_str() + colon + 1)) return false;
  std::string sym = s.substr(0, colon);
  int kth = 0;
  const size_t hash = sym.find('#');
  if (hash != std::string::npos) { kth = atoi(sym.c_str() + hash + 1); sym = sym.substr(0, hash); }
  if (sym != name) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  const int n = ++calls[sym];
  return kth == 0 || n == kth;
}

bool Identify(int fd, uint64_t* ino, uint64_t* dev, size_t* size) {
  struct stat st;
  if (fstat(fd, &st) != 0) return false;
  *ino = static_cast<uint64_t>(st.st_ino);
  *dev = static_cast<uint64_t>(st.st_dev);
  *size = static_cast<size_t>(st.st_size);
  return true;
}

CUresult MemCreate(CUmemGenericAllocationHandle* out, size_t bytes, const CUmemAllocationProp* prop, unsigned long long) {
  if (Inject("cuMemCreate")) return CUDA_ERROR_UNKNOWN;
  if (bytes == 0 || bytes % kGranularity != 0) return CUDA_ERROR_INVALID_VALUE;
  const int fd = memfd_create("mxkvsim_vmm", MFD_CLOEXEC);
  if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes)) != 0) { if (fd >= 0) close(fd); return CUDA_ERROR_OUT_OF_MEMORY; }
  Handle h;
  h.fd = fd; h.bytes = bytes; h.device = prop->location.id;
  size_t sz;
  Identify(fd, &h.ino, &h.dev, &sz);
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles[g_next] = h;
  *out = g_next++;
  return CUDA_SUCCESS;
}

CUresult MemRelease(CUmemGenericAllocationHandle h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(h);
  if (it == g_handles.end()) return CUDA_ERROR_INVALID_VALUE;
  if (it->second.fd >= 0) close(it->second.fd);   // mappings made from it stay valid (as on the device)
  g_handles.erase(it);
  return CUDA_SUCCESS;
}

CUresult MemExport(void* out, CUmemGenericAllocationHandle h, CUmemAllocationHandleType type, unsigned long long) {
  if (Inject("cuMemExportToShareableHandle")) return CUDA_ERROR_UNKNOWN;
  if (type != CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) return CUDA_ERROR_NOT_SUPPORTED;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(h);
  if (it == g_handles.end()) return CUDA_ERROR_INVALID_VALUE;
  const int fd = dup(it->second.fd);
  if (fd < 0) return CUDA_ERROR_UNKNOWN;
  *static_cast<int*>(out) = fd;
  return CUDA_SUCCESS;
}

McCtrl* MapCtrl(int fd) {
  void* p = mmap(nullptr, kCtrlBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  return p == MAP_FAILED ? nullptr : static_cast<McCtrl*>(p);
}

CUresult MemImport(CUmemGenericAllocationHandle* out, void* os_handle, CUmemAllocationHandleType type) {
  if (Inject("cuMemImportFromShareableHandle")) return CUDA_ERROR_UNKNOWN;
  if (type != CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) return CUDA_ERROR_NOT_SUPPORTED;
  const int theirs = static_cast<int>(reinterpret_cast<intptr_t>(os_handle));
  Handle h;
  size_t size = 0;
  if (!Identify(theirs, &h.ino, &h.dev, &size)) return CUDA_ERROR_INVALID_VALUE;
  h.fd = dup(theirs);                              // the caller closes its descriptor afterwards
  if (h.fd < 0) return CUDA_ERROR_UNKNOWN;
  h.bytes = size;
  if (size == kCtrlBytes) {
    h.ctrl = MapCtrl(h.fd);
    if (h.ctrl == nullptr || h.ctrl->magic != kCtrlMagic) { close(h.fd); return CUDA_ERROR_INVALID_VALUE; }
    h.multicast = true;
    h.bytes = h.ctrl->bytes;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles[g_next] = h;
  *out = g_next++;
  return CUDA_SUCCESS;
}

CUresult MemGetGranularity(size_t* g, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) {
  *g = kGranularity;
  return CUDA_SUCCESS;
}

CUresult MemAddressReserve(CUdeviceptr* out, size_t size, size_t, CUdeviceptr, unsigned long long) {
  if (Inject("cuMemAddressReserve")) return CUDA_ERROR_UNKNOWN;
  void* p = mmap(nullptr, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return CUDA_ERROR_OUT_OF_MEMORY;
  std::lock_guard<std::mutex> lk(g_mu);
  g_reserved[reinterpret_cast<uintptr_t>(p)] = size;
  *out = reinterpret_cast<CUdeviceptr>(p);
  return CUDA_SUCCESS;
}

CUresult MemAddressFree(CUdeviceptr va, size_t size) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_reserved.erase(static_cast<uintptr_t>(va));
  munmap(reinterpret_cast<void*>(va), size);
  return CUDA_SUCCESS;
}

CUresult MemMap(CUdeviceptr va, size_t size, size_t offset, CUmemGenericAllocationHandle h, unsigned long long) {
  if (Inject("cuMemMap")) return CUDA_ERROR_UNKNOWN;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(h);
  if (it == g_handles.end() || offset != 0 || size > it->second.bytes) return CUDA_ERROR_INVALID_VALUE;
  Mapping m;
  m.va = static_cast<uintptr_t>(va); m.bytes = size; m.multicast = it->second.multicast;
  m.ino = it->second.ino; m.dev = it->second.dev; m.ctrl = it->second.ctrl;
  if (!m.multicast) {
    void* p = mmap(reinterpret_cast<void*>(va), size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, it->second.fd, 0);
    if (p == MAP_FAILED) return CUDA_ERROR_UNKNOWN;
  }                                              // (a multicast range stays PROT_NONE: only multimem may touch it)
  g_maps.push_back(m);
  return CUDA_SUCCESS;
}

CUresult MemUnmap(CUdeviceptr va, size_t size) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < g_maps.size(); ++i) {
    if (g_maps[i].va != static_cast<uintptr_t>(va)) continue;
    // back to an inaccessible reservation
    mmap(reinterpret_cast<void*>(va), size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_FIXED, -1, 0);
    g_maps.erase(g_maps.begin() + static_cast<long>(i));
    return CUDA_SUCCESS;
  }
  return CUDA_ERROR_INVALID_VALUE;
}

CUresult MemSetAccess(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) { return CUDA_SUCCESS; }

CUresult MulticastCreate(CUmemGenericAllocationHandle* out, const CUmulticastObjectProp* prop) {
  if (Inject("cuMulticastCreate")) return CUDA_ERROR_UNKNOWN;
  if (prop->numDevices < 1 || prop->numDevices > kMaxDev || prop->size % kGranularity != 0) return CUDA_ERROR_INVALID_VALUE;
  const int fd = memfd_create("mxkvsim_mc", MFD_CLOEXEC);
  if (fd < 0 || ftruncate(fd, kCtrlBytes) != 0) { if (fd >= 0) close(fd); return CUDA_ERROR_OUT_OF_MEMORY; }
  Handle h;
  h.multicast = true; h.fd = fd; h.bytes = prop->size;
  h.ctrl = MapCtrl(fd);
  if (h.ctrl == nullptr) { close(fd); return CUDA_ERROR_UNKNOWN; }
  std::memset(h.ctrl, 0, kCtrlBytes);
  h.ctrl->num_devices = prop->numDevices;
  h.ctrl->bytes = prop->size;
  __atomic_store_n(&h.ctrl->magic, kCtrlMagic, __ATOMIC_RELEASE);
  size_t sz;
  Identify(fd, &h.ino, &h.dev, &sz);
  std::lock_guard<std::mutex> lk(g_mu);
  g_handles[g_next] = h;
  *out = g_next++;
  return CUDA_SUCCESS;
}

CUresult MulticastAddDevice(CUmemGenericAllocationHandle mc, CUdevice dev) {
  if (Inject("cuMulticastAddDevice")) return CUDA_ERROR_UNKNOWN;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(mc);
  if (it == g_handles.end() || !it->second.multicast || dev < 0 || dev >= kMaxDev) return CUDA_ERROR_INVALID_VALUE;
  __atomic_fetch_or(&it->second.ctrl->added, 1u << dev, __ATOMIC_ACQ_REL);
  return CUDA_SUCCESS;
}

CUresult MulticastBindMem(CUmemGenericAllocationHandle mc, size_t mc_offset, CUmemGenericAllocationHandle mem, size_t mem_offset,
                          size_t size, unsigned long long) {
  if (Inject("cuMulticastBindMem")) return CUDA_ERROR_UNKNOWN;
  std::lock_guard<std::mutex> lk(g_mu);
  auto im = g_handles.find(mc);
  auto ia = g_handles.find(mem);
  if (im == g_handles.end() || ia == g_handles.end() || !im->second.multicast || ia->second.multicast) return CUDA_ERROR_INVALID_VALUE;
  const int d = ia->second.device;
  McCtrl* c = im->second.ctrl;
  if (d < 0 || d >= kMaxDev || !(c->added & (1u << d))) return CUDA_ERROR_INVALID_DEVICE;     // AddDevice comes first
  if (mc_offset + size > c->bytes || mem_offset + size > ia->second.bytes) return CUDA_ERROR_INVALID_VALUE;
  McCtrl::Slot& s = c->slot[d];
  s.ino = ia->second.ino; s.dev = ia->second.dev; s.mem_offset = mem_offset; s.mc_offset = mc_offset; s.size = size;
  __atomic_store_n(&s.bound, 1u, __ATOMIC_RELEASE);
  return CUDA_SUCCESS;
}

CUresult MulticastGetGranularity(size_t* g, const CUmulticastObjectProp*, CUmulticastGranularity_flags) {
  *g = kGranularity;
  return CUDA_SUCCESS;
}

CUresult DeviceGet(CUdevice* out, int ordinal) { *out = ordinal; return CUDA_SUCCESS; }
CUresult DeviceGetAttribute(int* v, CUdevice_attribute attr, CUdevice) {
  *v = attr == CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED ? 1 : 0;
  return CUDA_SUCCESS;
}

}  // namespace

// cudaGetDriverEntryPoint of the stand-in runtime
void* DriverEntry(const char* name) {
  static const bool off = getenv("MXKV_SIM_NO_VMM") != nullptr;
  if (off) return nullptr;
  const std::string n = name;
#define ENTRY(sym, fn) if (n == sym) return reinterpret_cast<void*>(&fn)
  ENTRY("cuMemCreate", MemCreate); ENTRY("cuMemRelease", MemRelease);
  ENTRY("cuMemExportToShareableHandle", MemExport); ENTRY("cuMemImportFromShareableHandle", MemImport);
  ENTRY("cuMemGetAllocationGranularity", MemGetGranularity);
  ENTRY("cuMemAddressReserve", MemAddressReserve); ENTRY("cuMemAddressFree", MemAddressFree);
  ENTRY("cuMemMap", MemMap); ENTRY("cuMemUnmap", MemUnmap); ENTRY("cuMemSetAccess", MemSetAccess);
  ENTRY("cuMulticastCreate", MulticastCreate); ENTRY("cuMulticastAddDevice", MulticastAddDevice);
  ENTRY("cuMulticastBindMem", MulticastBindMem); ENTRY("cuMulticastGetGranularity", MulticastGetGranularity);
  ENTRY("cuDeviceGet", DeviceGet); ENTRY("cuDeviceGetAttribute", DeviceGetAttribute);
#undef ENTRY
  return nullptr;
}

}  // namespace sim

// A multicast address -> the local mappings of the allocations bound to the object, in device order (host_emu.h's
// multimem.ld_reduce / multimem.st).  Returns how many; aborts on an address that is no multicast address or whose
// object has a device bound that this process has not mapped.
extern "C" int mxkv_sim_multimem(const void* p, void** out) {
  using namespace sim;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  // (one lookup per multicast range and thread: a kernel makes millions of these accesses)
  struct Cache { uintptr_t begin = 0, end = 0; int n = 0; uintptr_t base[kMaxDev]; };
  static thread_local Cache cache;
  if (a >= cache.begin && a < cache.end) {
    for (int i = 0; i < cache.n; ++i) out[i] = reinterpret_cast<void*>(cache.base[i] + (a - cache.begin));
    return cache.n;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  for (const Mapping& m : g_maps) {
    if (!m.multicast || a < m.va || a >= m.va + m.bytes) continue;
    const uint64_t off = a - m.va;
    int n = 0;
    for (int d = 0; d < kMaxDev; ++d) {
      const McCtrl::Slot& s = m.ctrl->slot[d];
      if (!__atomic_load_n(&s.bound, __ATOMIC_ACQUIRE)) continue;
      if (off < s.mc_offset || off >= s.mc_offset + s.size) continue;
      const Mapping* local = nullptr;
      for (const Mapping& q : g_maps)
        if (!q.multicast && q.ino == s.ino && q.dev == s.dev) { local = &q; break; }
      if (local == nullptr) { fprintf(stderr, "[mxkv sim] multicast object names an allocation this process has not mapped\n"); abort(); }
      out[n++] = reinterpret_cast<void*>(local->va + s.mem_offset + (off - s.mc_offset));
    }
    if (n != static_cast<int>(m.ctrl->num_devices)) {
      fprintf(stderr, "[mxkv sim] multimem access while %d of %u devices have bound memory\n", n, m.ctrl->num_devices);
      abort();
    }
    // the arena binds whole segments at offset 0: remember the range
    bool whole = true;
    for (int d = 0; d < kMaxDev; ++d)
      if (m.ctrl->slot[d].bound && (m.ctrl->slot[d].mc_offset != 0 || m.ctrl->slot[d].size != m.bytes)) whole = false;
    if (whole) {
      cache.begin = m.va; cache.end = m.va + m.bytes; cache.n = n;
      for (int i = 0; i < n; ++i) cache.base[i] = reinterpret_cast<uintptr_t>(out[i]) - off;
    }
    return n;
  }
  fprintf(stderr, "[mxkv sim] multimem instruction on %p, which is not a multicast address\n", p);
  abort();
}
