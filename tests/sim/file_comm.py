"""A process group made of files: the only collective the engine's bootstrap needs is an all-gather of small
byte blobs, and for the multi-process simulator tests (no torch next to the stand-in libcudart) a shared
directory is enough.  Test infrastructure."""
import os
import time


class FileComm(object):
    def __init__(self, rank, world, directory, timeout=600.0):
        self.rank, self.world, self.dir, self.timeout = rank, world, directory, timeout
        self.seq = 0

    def allgather(self, payload):
        self.seq += 1
        mine = os.path.join(self.dir, "%06d_%d" % (self.seq, self.rank))
        with open(mine + ".tmp", "wb") as f:
            f.write(bytes(payload))
        os.rename(mine + ".tmp", mine)                      # atomic publication
        out = []
        deadline = time.time() + self.timeout
        for r in range(self.world):
            path = os.path.join(self.dir, "%06d_%d" % (self.seq, r))
            while not os.path.exists(path):
                if time.time() > deadline:
                    raise RuntimeError("rank %d: all-gather %d timed out waiting for rank %d" % (self.rank, self.seq, r))
                time.sleep(0.0005)
            with open(path, "rb") as f:
                out.append(f.read())
        return out

    def barrier(self):
        self.allgather(b"b")

    def allgather_int(self, x):
        return [int.from_bytes(b, "little", signed=True) for b in self.allgather(int(x).to_bytes(16, "little", signed=True))]
