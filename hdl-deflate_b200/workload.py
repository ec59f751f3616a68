"""CPU definition of the synthetic "random+repeat" block workload (BASELINE config 2,
SURVEY.md 8(d)).  Bit-identical to the device generator csrc/hdlz_workload.cu; used by
the tests to check that generator and to make small seeded inputs."""

DEFAULT_SEED = 0xDEF1A7E
_M = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15


def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def block(index, length=2048, seed=DEFAULT_SEED):
    """Bytes of block `index`."""
    s = (seed ^ ((index * 0xD1342543DE82EF95) & _M)) & _M
    out = bytearray()
    while len(out) < length:
        s = (s + _GOLD) & _M
        r = _mix64(s)
        pos = len(out)
        if (r & 1) == 0 or pos == 0:
            run = 1 + ((r >> 1) & 7)
            s = (s + _GOLD) & _M
            b = _mix64(s)
            for k in range(run):
                out.append((b >> (8 * k)) & 255)
        else:
            m = 3 + ((r >> 1) % 10)
            d = 1 + ((r >> 8) % min(32, pos))
            for _ in range(m):
                out.append(out[-d])
    return bytes(out[:length])


def blocks(first, count, length=2048, seed=DEFAULT_SEED):
    return [block(first + i, length, seed) for i in range(count)]
