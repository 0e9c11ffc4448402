"""numpy restatement of the engine's counter-based index stream (Philox4x32-10, Salmon et al. 2011;
flexpart_b200/csrc/fpb_kernels.cu philox4x32_10 / Rng::uniform): the uniform that replaces ran3 in
`nrand=int(ran3(idummy)*real(maxrand-1))+1` (src/advance.f90:153, src/initialize.f90:68) in the
production RNG modes.  Test infrastructure: lets the CPU oracle consume exactly the indices the
device draws, so that the benchmarked RNG mode can be compared with it particle by particle."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, np.uint32).copy() for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def index_uniform(cfg, slots, itime, stream):
    """Rng::uniform(stream) of the particles in `slots` at time `itime` (float32)."""
    pid = (np.int64(cfg.part_id_offset) + np.int64(cfg.part_id_stride or 1) * np.asarray(slots, np.int64)) & 0xFFFFFFFF
    seed = int(cfg.seed)
    x, _, _, _ = philox4x32_10(pid.astype(np.uint32), np.uint32(itime & 0xFFFFFFFF), np.uint32(stream), np.uint32(0),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def index_queue(cfg, p, itime):
    """The uniforms the oracle's sequential loop consumes at `itime`, in particle order: for every
    active particle one for initialize (stream 1) when it is new, then one for advance (stream 2)."""
    n = p.numpart
    slots = np.arange(n)
    act = p.itra1[:n] == itime
    new = act & ((p.itramem[:n] == itime) | (itime == 0))
    u1, u2 = index_uniform(cfg, slots, itime, 1), index_uniform(cfg, slots, itime, 2)
    q = np.empty((n, 2), np.float32)
    q[:, 0], q[:, 1] = u1, u2
    keep = np.stack([new, act], axis=1)
    return np.ascontiguousarray(q[keep])
