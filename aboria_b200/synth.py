"""Counter-based synthetic inputs: any rank can generate any
particle.  u(seed, k) = (splitmix64(seed ^ k) >> 11) * 2^-53 in [0, 1)."""
import numpy as np

SEED = 20260101
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def uniform01(seed, counters):
    with np.errstate(over="ignore"):
        c = np.asarray(counters, dtype=np.uint64)
        bits = splitmix64(np.uint64(seed) ^ c)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_positions(n, D, low=0.0, high=1.0, seed=SEED, first_id=0):
    """positions of particles first_id .. first_id+n-1, uniform in [low, high)^D"""
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    k = ids[:, None] * np.uint64(D) + np.arange(D, dtype=np.uint64)[None, :]
    u = uniform01(seed, k)
    low = np.broadcast_to(np.asarray(low, dtype=np.float64), (D,))
    high = np.broadcast_to(np.asarray(high, dtype=np.float64), (D,))
    p = low + u * (high - low)
    return np.minimum(p, np.nextafter(high, low))  # keep strictly below `high`


def vector(n, seed=SEED + 1, first_id=0, width=1):
    ids = np.arange(first_id * width, (first_id + n) * width, dtype=np.uint64)
    return uniform01(seed, ids)


def clustered_positions(n, D=3, n_blobs=64, sigma=0.03, background=0.1, seed=SEED, periodic_dims=(True, True, False), first_id=0):
    """Gaussian blobs + uniform background in [0,1)^D; coordinates
    are wrapped in periodic dims and reflected in the others.  Particles first_id .. first_id+n-1."""
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    u = uniform01(seed + 7, ids)
    centres = uniform01(seed + 11, np.arange(n_blobs * D, dtype=np.uint64)).reshape(n_blobs, D)
    blob = (uniform01(seed + 13, ids) * n_blobs).astype(np.int64) % n_blobs
    k = ids[:, None] * np.uint64(2 * D) + np.arange(2 * D, dtype=np.uint64)[None, :]
    uu = uniform01(seed + 17, k)
    # Box-Muller
    g = np.sqrt(-2.0 * np.log(1.0 - uu[:, :D])) * np.cos(2.0 * np.pi * uu[:, D:])
    p = centres[blob] + sigma * g
    bg = uniform_positions(n, D, seed=seed + 19, first_id=first_id)
    p = np.where((u < background)[:, None], bg, p)
    for d in range(D):
        if periodic_dims[d] if d < len(periodic_dims) else False:
            p[:, d] = p[:, d] - np.floor(p[:, d])
        else:
            q = np.mod(p[:, d], 2.0)
            p[:, d] = np.where(q >= 1.0, 2.0 - q, q)
    return np.clip(p, 0.0, np.nextafter(1.0, 0.0))


def torch_uniform_positions(n, D, low, high, seed, first_id, device):
    """same numbers as uniform_positions, generated on the device (int64
    arithmetic wraps mod 2^64; logical shifts emulated by masking)."""
    import torch

    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)

    def to_i64(v):
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= (1 << 63) else v

    ids = torch.arange(first_id, first_id + n, dtype=torch.int64, device=device)
    k = ids[:, None] * D + torch.arange(D, dtype=torch.int64, device=device)[None, :]
    x = k ^ to_i64(int(seed))
    x = x + to_i64(0x9E3779B97F4A7C15)
    x = (x ^ lsr(x, 30)) * to_i64(0xBF58476D1CE4E5B9)
    x = (x ^ lsr(x, 27)) * to_i64(0x94D049BB133111EB)
    x = x ^ lsr(x, 31)
    u = lsr(x, 11).to(torch.float64) * (1.0 / 9007199254740992.0)
    low_t = torch.as_tensor(np.broadcast_to(np.asarray(low, dtype=np.float64), (D,)).copy(), device=device)
    high_t = torch.as_tensor(np.broadcast_to(np.asarray(high, dtype=np.float64), (D,)).copy(), device=device)
    p = low_t + u * (high_t - low_t)
    lim = torch.as_tensor(np.nextafter(high_t.cpu().numpy(), low_t.cpu().numpy()), device=device)
    return torch.minimum(p, lim).contiguous()
