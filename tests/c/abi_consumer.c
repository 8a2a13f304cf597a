/* abi_consumer.c -- a plain C consumer of include/mxkv_b200.h: proves that the header is valid C99,
 * that the library can be linked from C, and that the host-only part of the contract works without a
 * GPU: handles, host NDArrays, key-type rules, error reporting through MXGetLastError (0 / -1, never an
 * exception across the boundary, include/mxnet/c_api_error.h:40-58).  Built and run by
 * tests/test_abi.py::test_plain_c_consumer.  Exit code 0 = all checks passed. */
#include <stdio.h>
#include <string.h>
#include "mxkv_b200.h"

#define CHECK(cond, what)                                                                  \
  do {                                                                                     \
    if (!(cond)) { fprintf(stderr, "FAILED: %s (%s)\n", what, MXGetLastError()); return 1; } \
  } while (0)

int main(void) {
  KVStoreHandle kv = NULL;
  const char* type = NULL;
  int rank = -1, size = -1, version = 0, dim = 0, dtype = -1, dev_type = 0, dev_id = 0, stype = -1;
  const int64_t* pshape = NULL;
  int64_t shape[2] = {3, 4};
  float host[12], back[12];
  NDArrayHandle a = NULL, out = NULL, bad = NULL;
  int i, keys[1] = {7};
  const char* skeys[1] = {"seven"};

  CHECK(MXGetVersion(&version) == 0 && version > 0, "MXGetVersion");
  CHECK(MXKVStoreCreate("local", &kv) == 0 && kv != NULL, "MXKVStoreCreate");
  CHECK(MXKVStoreGetType(kv, &type) == 0 && strcmp(type, "local") == 0, "MXKVStoreGetType");
  CHECK(MXKVStoreGetRank(kv, &rank) == 0 && rank == 0, "MXKVStoreGetRank");
  CHECK(MXKVStoreGetGroupSize(kv, &size) == 0 && size == 1, "MXKVStoreGetGroupSize");

  /* a host NDArray round trip */
  for (i = 0; i < 12; ++i) host[i] = (float)i * 0.5f;
  CHECK(MXNDArrayCreate64(shape, 2, 1 /* cpu */, 0, 0, 0 /* float32 */, &a) == 0, "MXNDArrayCreate64");
  CHECK(MXNDArraySyncCopyFromCPU(a, host, 12) == 0, "MXNDArraySyncCopyFromCPU");
  CHECK(MXNDArrayGetShape64(a, &dim, &pshape) == 0 && dim == 2 && pshape[0] == 3 && pshape[1] == 4, "GetShape64");
  CHECK(MXNDArrayGetDType(a, &dtype) == 0 && dtype == 0, "MXNDArrayGetDType");
  CHECK(MXNDArrayGetContext(a, &dev_type, &dev_id) == 0 && dev_type == 1, "MXNDArrayGetContext");
  CHECK(MXNDArrayGetStorageType(a, &stype) == 0 && stype == 0, "MXNDArrayGetStorageType");

  /* init parks the value in the store; a host pull of a never-pushed key needs no GPU */
  CHECK(MXKVStoreInit(kv, 1, keys, &a) == 0, "MXKVStoreInit");
  CHECK(MXNDArrayCreate64(shape, 2, 1, 0, 0, 0, &out) == 0, "MXNDArrayCreate64(out)");
  CHECK(MXKVStorePull(kv, 1, keys, &out, 0) == 0, "MXKVStorePull");
  CHECK(MXNDArraySyncCopyToCPU(out, back, 12) == 0 && memcmp(host, back, sizeof(host)) == 0, "pulled value");

  /* error contract: -1 plus a message, the process keeps running */
  CHECK(MXKVStoreInit(kv, 1, keys, &a) == -1 && strlen(MXGetLastError()) > 0, "duplicate init is rejected");
  CHECK(MXKVStoreInitEx(kv, 1, skeys, &a) == -1 && strstr(MXGetLastError(), "Mixed key types") != NULL,
        "int and str keys cannot be mixed (kvstore_local.h:344-347)");
  CHECK(MXNDArrayCreate64(shape, 2, 1, 0, 0, 99 /* no such dtype */, &bad) == -1, "bad dtype is rejected");
  CHECK(MXKVStoreCreate("dist_sync", &bad) == -1, "distributed types are out of scope");

  CHECK(MXNDArrayFree(a) == 0 && MXNDArrayFree(out) == 0, "MXNDArrayFree");
  CHECK(MXKVStoreFree(kv) == 0, "MXKVStoreFree");
  printf("abi_consumer ok (library version %d)\n", version);
  return 0;
}
