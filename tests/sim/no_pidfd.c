/* no_pidfd.c -- LD_PRELOAD shim for one test: pidfd_open(2) fails with EPERM, as under the ptrace policy of the
 * containers the hardware runs happen in, so that csrc/vmm_arena.cc takes its second descriptor transport (abstract
 * unix sockets + SCM_RIGHTS).  Every other system call goes through.  Test infrastructure. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <stdarg.h>
#include <sys/syscall.h>
#include <unistd.h>

#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif

long syscall(long number, ...) {
  static long (*real)(long, ...) = 0;
  if (!real) real = (long (*)(long, ...))dlsym(RTLD_NEXT, "syscall");
  if (number == SYS_pidfd_open) { errno = EPERM; return -1; }
  va_list ap;
  va_start(ap, number);
  long a = va_arg(ap, long), b = va_arg(ap, long), c = va_arg(ap, long), d = va_arg(ap, long), e = va_arg(ap, long),
       f = va_arg(ap, long);
  va_end(ap);
  return real(number, a, b, c, d, e, f);
}
