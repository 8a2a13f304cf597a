// vmm_arena.cc -- engine-owned NVSwitch multicast memory for the one-process-per-GPU arena.
//
// Round 1 took multicast-capable arrays from the embedding framework (torch.distributed._symmetric_memory).
// Here the engine builds them itself, so that every array of the peer-mapped arena -- and therefore any consumer
// of the C ABI, torch or not -- can take the NVLS kernel (multimem.ld_reduce / multimem.st):
//   * a segment is a CUDA VMM allocation (cuMemCreate) per rank, exported as a POSIX file descriptor, duplicated
//     into every peer with pidfd_getfd(2) -- or, where the container forbids that, passed over an abstract unix
//     socket with SCM_RIGHTS -- (the bootstrap all-gather carries pid + fd number, nothing else), imported
//     and mapped there (cuMemImportFromShareableHandle / cuMemMap / cuMemSetAccess): the peer pointers that
//     cudaIpcOpenMemHandle used to provide;
//   * rank 0 creates ONE multicast object per segment (cuMulticastCreate), every rank adds its GPU
//     (cuMulticastAddDevice) and binds its allocation (cuMulticastBindMem); mapping the object gives the multicast
//     alias: a load-reduce through it sums the n copies in the switch, a store through it lands in all of them.
// The driver API is reached through cudaGetDriverEntryPoint (the library keeps linking the CUDA runtime only, and
// still loads on a machine without a driver).  Whatever is missing -- multicast support, pidfd_getfd, a driver
// symbol -- makes the collective probe fail on every rank together, and the arena falls back to cudaMalloc +
// cudaIpc (peer pointers, no multicast alias).
#include "runtime.h"
#include <cuda.h>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <cstddef>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/uio.h>
#include <sys/un.h>
#include <unistd.h>

#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif

namespace mxkv {

namespace {

struct DriverApi {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemGetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  bool ok = false;
};

template <typename F>
bool Load(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

const DriverApi& Api() {
  static DriverApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.ok = Load("cuMemCreate", &a.MemCreate) && Load("cuMemRelease", &a.MemRelease) &&
           Load("cuMemExportToShareableHandle", &a.MemExport) && Load("cuMemImportFromShareableHandle", &a.MemImport) &&
           Load("cuMemGetAllocationGranularity", &a.MemGetGranularity) &&
           Load("cuMemAddressReserve", &a.MemAddressReserve) && Load("cuMemAddressFree", &a.MemAddressFree) &&
           Load("cuMemMap", &a.MemMap) && Load("cuMemUnmap", &a.MemUnmap) && Load("cuMemSetAccess", &a.MemSetAccess) &&
           Load("cuMulticastCreate", &a.MulticastCreate) && Load("cuMulticastAddDevice", &a.MulticastAddDevice) &&
           Load("cuMulticastBindMem", &a.MulticastBindMem) &&
           Load("cuMulticastGetGranularity", &a.MulticastGetGranularity) && Load("cuDeviceGet", &a.DeviceGet) &&
           Load("cuDeviceGetAttribute", &a.DeviceGetAttribute);
  }
  return a;
}

int DupFromPeer(int pid, int fd) {
  const int pidfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
  if (pidfd < 0) return -1;
  const int got = static_cast<int>(syscall(SYS_pidfd_getfd, pidfd, fd, 0));
  close(pidfd);
  return got;
}

// first failure of this process, for the one-line note on the fallback (MXKV_B200_ARENA_VMM_VERBOSE=1)
std::string g_why;
bool Ok(const char* what, CUresult r) {
  if (r == CUDA_SUCCESS) return true;
  if (g_why.empty()) g_why = std::string(what) + " -> CUresult " + std::to_string(static_cast<int>(r));
  return false;
}
bool Ok(const char* what, bool v) {
  if (!v && g_why.empty()) g_why = std::string(what) + " failed (errno " + std::to_string(errno) + ")";
  return v;
}

// ---- descriptor transport #2: abstract unix sockets + SCM_RIGHTS (what survives a locked-down ptrace policy) ----
void FdSockName(int pid, uint64_t tag, sockaddr_un* addr, socklen_t* len) {
  std::memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  const int n = std::snprintf(addr->sun_path + 1, sizeof(addr->sun_path) - 1, "mxkv_b200_%d_%llu", pid,
                              static_cast<unsigned long long>(tag));      // leading NUL: abstract namespace
  *len = static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}
int FdServerOpen(int pid, uint64_t tag) {
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return -1;
  sockaddr_un addr; socklen_t len;
  FdSockName(pid, tag, &addr, &len);
  if (bind(s, reinterpret_cast<sockaddr*>(&addr), len) != 0 || listen(s, 64) != 0) { close(s); return -1; }
  return s;
}
int FdClientConnect(int pid, uint64_t tag) {
  const int s = socket(AF_UNIX, SOCK_STREAM, 0);
  if (s < 0) return -1;
  sockaddr_un addr; socklen_t len;
  FdSockName(pid, tag, &addr, &len);
  if (connect(s, reinterpret_cast<sockaddr*>(&addr), len) != 0) { close(s); return -1; }
  return s;
}
bool FdServerSendOne(int srv, const int* fds, int n) {       // accept one peer, hand it n descriptors
  const int c = accept(srv, nullptr, nullptr);
  if (c < 0) return false;
  char byte = 'x';
  iovec iov{&byte, 1};
  alignas(cmsghdr) char ctl[CMSG_SPACE(2 * sizeof(int))];
  std::memset(ctl, 0, sizeof(ctl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctl; msg.msg_controllen = CMSG_SPACE(n * sizeof(int));
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(n * sizeof(int));
  std::memcpy(CMSG_DATA(cm), fds, n * sizeof(int));
  const bool ok = sendmsg(c, &msg, 0) == 1;
  close(c);
  return ok;
}
bool FdClientRecv(int c, int* fds, int n) {
  char byte = 0;
  iovec iov{&byte, 1};
  alignas(cmsghdr) char ctl[CMSG_SPACE(2 * sizeof(int))];
  std::memset(ctl, 0, sizeof(ctl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctl; msg.msg_controllen = sizeof(ctl);
  if (recvmsg(c, &msg, 0) != 1) return false;
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (cm == nullptr || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS ||
      cm->cmsg_len != CMSG_LEN(n * sizeof(int)))
    return false;
  std::memcpy(fds, CMSG_DATA(cm), n * sizeof(int));
  return true;
}

struct Msg {            // one rank's contribution to a step of the protocol
  int64_t ok;
  int64_t pid, fd_mem, fd_mc;
  uint64_t bytes;
};

}  // namespace

// Collective.  Returns true with s filled (base[r] = rank r's allocation as mapped here, mc = multicast alias) on
// EVERY rank, or false on every rank (nothing left mapped).  `bytes` is rounded up to the multicast granularity.
bool ProcessGroup::NewSegmentVmm(size_t min_bytes, Segment* out) {
  const DriverApi& a = Api();
  Segment s;
  std::memset(&s, 0, sizeof(s));
  CUmemGenericAllocationHandle h_mem[kMaxRanks] = {0};
  CUmemGenericAllocationHandle h_mc = 0;
  int fd_mem = -1, fd_mc = -1;
  size_t bytes = 0;
  bool ok = a.ok && world_ > 1;
  CUdevice cudev = 0;
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  CUmulticastObjectProp mcprop;
  std::memset(&mcprop, 0, sizeof(mcprop));

  auto agree = [&](bool mine, Msg* all_out = nullptr, int64_t pid = 0, int64_t f1 = -1, int64_t f2 = -1) -> bool {
    Msg m{mine ? 1 : 0, pid, f1, f2, bytes};
    std::vector<Msg> all(world_);
    AllGather(&m, sizeof(Msg), all.data());
    bool every = true;
    for (int r = 0; r < world_; ++r) every = every && all[r].ok == 1 && all[r].bytes == bytes;
    if (all_out) std::memcpy(all_out, all.data(), sizeof(Msg) * world_);
    return every;
  };

  // ---- step 1: capability, granularity, my allocation + its fd; rank 0: the multicast object + its fd
  if (ok) {
    int mc_ok = 0;
    ok = Ok("cuDeviceGet", a.DeviceGet(&cudev, dev_)) &&
         Ok("cuDeviceGetAttribute(MULTICAST_SUPPORTED)", a.DeviceGetAttribute(&mc_ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev)) &&
         Ok("device reports multicast support", mc_ok != 0);
  }
  if (ok) {
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev_;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mcprop.numDevices = static_cast<unsigned>(world_);
    mcprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g_mem = 0, g_mc = 0;
    mcprop.size = size_t(2) << 20;
    ok = Ok("cuMemGetAllocationGranularity", a.MemGetGranularity(&g_mem, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED)) &&
         Ok("cuMulticastGetGranularity", a.MulticastGetGranularity(&g_mc, &mcprop, CU_MULTICAST_GRANULARITY_RECOMMENDED)) &&
         g_mem > 0 && g_mc > 0;
    if (ok) {
      const size_t g = std::max(g_mem, g_mc);
      bytes = (min_bytes + g - 1) / g * g;
      mcprop.size = bytes;
    }
  }
  if (ok) ok = Ok("cuMemCreate", a.MemCreate(&h_mem[rank_], bytes, &prop, 0));
  if (ok) ok = Ok("cuMemExportToShareableHandle", a.MemExport(&fd_mem, h_mem[rank_], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  if (ok && rank_ == 0) {
    ok = Ok("cuMulticastCreate", a.MulticastCreate(&h_mc, &mcprop)) &&
         Ok("cuMemExportToShareableHandle(multicast)", a.MemExport(&fd_mc, h_mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  }
  std::vector<Msg> peers(world_);
  bool all_ok = agree(ok, peers.data(), static_cast<int64_t>(getpid()), fd_mem, fd_mc);

  // ---- step 2: get the peers' descriptors into this process (pidfd_getfd; where the container forbids it --
  // EPERM under a restrictive ptrace policy -- an abstract unix socket per rank and SCM_RIGHTS), then import them
  int got_mem[kMaxRanks];
  int got_mc = -1;
  for (int r = 0; r < kMaxRanks; ++r) got_mem[r] = -1;
  if (all_ok) {
    prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0);      // Yama ptrace_scope = 1: let the peers duplicate from us
    all_ok = agree(true);                                     // (everybody has opted in before anybody tries)
  }
  bool have_fds = false;
  if (all_ok) {
    bool mine = true;
    for (int r = 0; r < world_ && mine; ++r) {
      if (r == rank_) continue;
      got_mem[r] = DupFromPeer(static_cast<int>(peers[r].pid), static_cast<int>(peers[r].fd_mem));
      mine = got_mem[r] >= 0;
      if (mine && r == 0) { got_mc = DupFromPeer(static_cast<int>(peers[0].pid), static_cast<int>(peers[0].fd_mc)); mine = got_mc >= 0; }
    }
    const int err = errno;
    have_fds = agree(mine);
    if (!have_fds) {
      for (int r = 0; r < world_; ++r) if (got_mem[r] >= 0) { close(got_mem[r]); got_mem[r] = -1; }
      if (got_mc >= 0) { close(got_mc); got_mc = -1; }
      // second transport: every rank serves its descriptors over an abstract unix socket
      const uint64_t tag = static_cast<uint64_t>(segs_.size());
      const int srv = FdServerOpen(static_cast<int>(getpid()), tag);
      bool up = agree(srv >= 0);                              // all sockets are listening
      int conn[kMaxRanks];
      for (int r = 0; r < kMaxRanks; ++r) conn[r] = -1;
      mine = up;
      if (up) {
        for (int r = 0; r < world_ && mine; ++r) {            // connects complete out of the listen backlog
          if (r == rank_) continue;
          conn[r] = FdClientConnect(static_cast<int>(peers[r].pid), tag);
          mine = conn[r] >= 0;
        }
        const int fds[2] = {fd_mem, rank_ == 0 ? fd_mc : -1};
        for (int k = 0; k < world_ - 1 && mine; ++k) mine = FdServerSendOne(srv, fds, rank_ == 0 ? 2 : 1);
        for (int r = 0; r < world_ && mine; ++r) {
          if (r == rank_) continue;
          int in[2] = {-1, -1};
          mine = FdClientRecv(conn[r], in, r == 0 ? 2 : 1);
          got_mem[r] = in[0];
          if (r == 0) got_mc = in[1];
        }
      }
      for (int r = 0; r < world_; ++r) if (conn[r] >= 0) close(conn[r]);
      if (srv >= 0) close(srv);
      if (!mine && g_why.empty())
        g_why = "descriptor exchange failed: pidfd_getfd errno " + std::to_string(err) + ", unix socket errno " + std::to_string(errno);
      have_fds = agree(mine);
    }
    all_ok = have_fds;
  }
  if (all_ok) {
    for (int r = 0; r < world_ && ok; ++r) {
      if (r == rank_) continue;
      ok = Ok("cuMemImportFromShareableHandle", a.MemImport(&h_mem[r], reinterpret_cast<void*>(static_cast<intptr_t>(got_mem[r])),
                                                            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    }
    if (ok && rank_ != 0)
      ok = Ok("cuMemImportFromShareableHandle(multicast)", a.MemImport(&h_mc, reinterpret_cast<void*>(static_cast<intptr_t>(got_mc)),
                                                                       CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    all_ok = agree(ok);          // (the exporters keep their descriptors open until everybody has imported)
  }
  for (int r = 0; r < world_; ++r) if (got_mem[r] >= 0) close(got_mem[r]);
  if (got_mc >= 0) close(got_mc);
  if (fd_mem >= 0) close(fd_mem);
  if (fd_mc >= 0) close(fd_mc);

  // ---- step 3: every GPU joins the multicast object, then binds its allocation
  if (all_ok) { ok = Ok("cuMulticastAddDevice", a.MulticastAddDevice(h_mc, cudev)); all_ok = agree(ok); }
  if (all_ok) { ok = Ok("cuMulticastBindMem", a.MulticastBindMem(h_mc, 0, h_mem[rank_], 0, bytes, 0)); all_ok = agree(ok); }

  // ---- step 4: map everything for this GPU
  CUdeviceptr va[kMaxRanks] = {0};
  CUdeviceptr va_mc = 0;
  if (all_ok) {
    CUmemAccessDesc acc;
    std::memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = dev_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    for (int r = 0; r < world_ && ok; ++r) {
      ok = Ok("cuMemAddressReserve", a.MemAddressReserve(&va[r], bytes, 0, 0, 0)) &&
           Ok("cuMemMap", a.MemMap(va[r], bytes, 0, h_mem[r], 0)) &&
           Ok("cuMemSetAccess", a.MemSetAccess(va[r], bytes, &acc, 1));
    }
    if (ok) {
      ok = Ok("cuMemAddressReserve(multicast)", a.MemAddressReserve(&va_mc, bytes, 0, 0, 0)) &&
           Ok("cuMemMap(multicast)", a.MemMap(va_mc, bytes, 0, h_mc, 0)) &&
           Ok("cuMemSetAccess(multicast)", a.MemSetAccess(va_mc, bytes, &acc, 1));
    }
    all_ok = agree(ok);
  }
  if (!all_ok) {
    if (!a.ok && g_why.empty()) g_why = "a CUDA driver entry point is missing";
    if (std::getenv("MXKV_B200_ARENA_VMM_VERBOSE") != nullptr || rank_ == 0)
      std::fprintf(stderr, "[mxkv_b200] rank %d: no engine-owned multicast arena (%s); using cudaMalloc + cudaIpc\n", rank_,
                   g_why.empty() ? "another rank could not set it up" : g_why.c_str());
    // undo whatever this rank got as far as (mappings die with the address ranges; handles are reference counted)
    for (int r = 0; r < world_; ++r) {
      if (va[r]) { a.MemUnmap(va[r], bytes); a.MemAddressFree(va[r], bytes); }
      if (h_mem[r]) a.MemRelease(h_mem[r]);
    }
    if (va_mc) { a.MemUnmap(va_mc, bytes); a.MemAddressFree(va_mc, bytes); }
    if (h_mc) a.MemRelease(h_mc);
    return false;
  }
  // the mappings keep the memory alive; the handles themselves are no longer needed
  for (int r = 0; r < world_; ++r) a.MemRelease(h_mem[r]);
  a.MemRelease(h_mc);
  for (int r = 0; r < world_; ++r) s.base[r] = reinterpret_cast<char*>(va[r]);
  s.mc = reinterpret_cast<char*>(va_mc);
  s.bytes = bytes;
  s.used = 0;
  s.vmm = true;
  *out = s;
  return true;
}

void ProcessGroup::FreeSegmentVmm(Segment& s) {
  const DriverApi& a = Api();
  if (!a.ok) return;
  for (int r = 0; r < world_; ++r)
    if (s.base[r]) { a.MemUnmap(reinterpret_cast<CUdeviceptr>(s.base[r]), s.bytes); a.MemAddressFree(reinterpret_cast<CUdeviceptr>(s.base[r]), s.bytes); }
  if (s.mc) { a.MemUnmap(reinterpret_cast<CUdeviceptr>(s.mc), s.bytes); a.MemAddressFree(reinterpret_cast<CUdeviceptr>(s.mc), s.bytes); }
}

}  // namespace mxkv
