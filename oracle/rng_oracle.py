"""TEST INFRASTRUCTURE ONLY: numpy restatement, bit for bit, of the device-side counter-based noise generator
(csrc/rng.cuh: Philox4x32-10 + an IEEE-only Box-Muller transform).

The reference draws its noise with tf.random_normal / tf.random_uniform (air_model.py:123-128, vae.py:23,37,
concrete.py:23) -- TensorFlow's Philox streams, which are not reproducible outside TensorFlow; parity runs inject
noise instead.  What is pinned here is that the in-kernel generator of the product is exactly this function of
(seed, step counter, stream, element index): every float operation below is a single correctly rounded IEEE fp32
operation (+, -, *, /, sqrt), which the kernels spell as __fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn, so numpy
float32 reproduces the bits."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
STREAM_SCALE, STREAM_SHIFT, STREAM_LATENT, STREAM_CONCRETE, STREAM_LIKE = 1, 2, 3, 4, 5
f32 = np.float32


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 arrays / scalars -> four uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF for v in (c0, c1, c2, c3)]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = c[0] * M0, c[2] * M1
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in c]


def uniform01(w):
    """[0, 1): the top 24 bits of a word."""
    return (w >> np.uint32(8)).astype(np.float32) * f32(2.0 ** -24)


def normal_pair(w0, w1):
    """Box-Muller from two words: radius from w0, angle from w1; only correctly rounded fp32 +, -, *, /, sqrt."""
    j = (w0 >> np.uint32(8)).astype(np.int64) + 1                     # 1 .. 2^24;  u1 = j * 2^-24 in (0, 1]
    e = np.floor(np.log2(j.astype(np.float64))).astype(np.int64)      # 31 - clz(j): exact for j <= 2^24
    m = (j.astype(np.float64) / (2.0 ** e)).astype(np.float32)        # in [1, 2), exact
    s = (m - f32(1)) / (m + f32(1))
    s2 = s * s
    p = f32(1.0 / 11.0)
    for c in (1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, 1.0):
        p = p * s2 + f32(c)
    ln_m = (f32(2) * s) * p
    ln_u = (e - 24).astype(np.float32) * f32(0.6931471805599453) + ln_m
    r = np.sqrt(f32(-2) * ln_u)
    k = (w1 >> np.uint32(8)).astype(np.int64)                         # angle = 2 pi k / 2^24 = (q + t) pi / 4
    q, frac = k >> 21, k & 0x1FFFFF
    t = frac.astype(np.float32) * f32(2.0 ** -21)
    odd = (q & 1) == 1
    x = np.where(odd, f32(1) - t, t).astype(np.float32) * f32(0.7853981633974483)
    x2 = x * x
    ps = f32(1.0 / 362880.0)
    for c in (-1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0):
        ps = ps * x2 + f32(c)
    sx = x * ps
    pc = f32(-1.0 / 3628800.0)
    for c in (1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5, 1.0):
        pc = pc * x2 + f32(c)
    cx = pc
    sx = np.where(odd, -sx, sx)                                       # theta = base * pi/2 + (+x | -x)
    base = ((q + 1) >> 1) & 3
    cos_t = np.select([base == 0, base == 1, base == 2], [cx, -sx, -cx], sx).astype(np.float32)
    sin_t = np.select([base == 0, base == 1, base == 2], [sx, cx, -sx], -cx).astype(np.float32)
    return (r * cos_t).astype(np.float32), (r * sin_t).astype(np.float32)


def normal_pair_fast(w0, w1):
    """The SFU form of the same transform that the kernels use for the likelihood stream (rng_normal_pair_fast): the
    device evaluates lg2 / rsqrt / sin / cos with the hardware approximations, so this agrees to ~1e-5, not bit for bit."""
    u1 = ((w0 >> np.uint32(8)).astype(np.float32) + f32(1)) * f32(2.0 ** -24)
    r = np.sqrt(f32(-1.3862943611198906) * np.log2(u1), dtype=np.float32)
    a = ((w1 >> np.uint32(8)).astype(np.float32) - f32(8388608.0)) * f32(3.7450702829238413e-07)
    return (r * np.cos(a, dtype=np.float32)).astype(np.float32), (r * np.sin(a, dtype=np.float32)).astype(np.float32)


def words(seed, counter, stream, n):
    """the 4 * ceil(n / 4) words of a stream: element i comes from block i >> 2, word i & 3"""
    blocks = np.arange((n + 3) // 4, dtype=np.uint64)
    return philox4x32(blocks & 0xFFFFFFFF, blocks >> 32, stream, counter & 0xFFFFFFFF, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def normals(seed, counter, stream, n):
    w = words(seed, counter, stream, n)
    pair = normal_pair_fast if stream == STREAM_LIKE else normal_pair
    a0, a1 = pair(w[0], w[1])
    a2, a3 = pair(w[2], w[3])
    return np.stack([a0, a1, a2, a3], axis=1).reshape(-1)[:n]


def uniforms(seed, counter, stream, n):
    w = words(seed, counter, stream, n)
    return np.stack([uniform01(v) for v in w], axis=1).reshape(-1)[:n]
