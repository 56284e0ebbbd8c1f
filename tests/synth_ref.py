"""CPU restatement (numpy) of the deterministic row generator in filodb_b200/csrc/synth_kernels.cu, used by the tests to rebuild
the rows the GPU generator encodes, so that its chunks can be compared byte-for-byte with the oracle's encoders."""
import numpy as np

M64 = (1 << 64) - 1
NOISE_SCALE = 1.0 / 37837.22772881784


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def series_key(seed, gid):
    return splitmix64(seed ^ ((gid * 0xD1342543DE82EF95) & M64))


def row_hash(key, row, salt):
    return splitmix64((key + ((row & 0xffffffff) << 3) + salt) & M64)


def group_id(seed, gid, n_groups):
    return splitmix64(seed ^ 0xA5A5A5A5 ^ ((gid * 0x9E3779B97F4A7C15) & M64)) % n_groups if n_groups > 0 else 0


def gen_series(seed, gid, rows, rows_per_chunk, t0, interval, jitter, value_kind, reset_period, nan_ppm, sin_table):
    key = series_key(seed, gid)
    ts = np.zeros(rows, np.int64)
    vals = np.zeros(rows, np.float64)
    v = 0.0
    for r in range(rows):
        t = t0 + r * interval
        if jitter > 0:
            t += int(row_hash(key, r, 3) % (2 * jitter + 1)) - jitter
        ts[r] = t
        last_in_chunk = ((r % rows_per_chunk) == rows_per_chunk - 1) or r == rows - 1
        if last_in_chunk and nan_ppm > 0 and (row_hash(key, r, 1) % 1000000) < nan_ppm:
            vals[r] = np.nan
            continue
        h = row_hash(key, r, 0)
        x = (h & 0xffff) + ((h >> 16) & 0xffff) + ((h >> 32) & 0xffff) + (h >> 48)
        noise = np.float64(float(int(x) - 131070)) * np.float64(NOISE_SCALE)
        s = (np.float64(15.0) + np.float64(sin_table[r])) + noise
        if value_kind == 0:
            vals[r] = s
            continue
        inc = s if s > 0.0 else np.float64(0.0)
        if value_kind == 2:
            inc = np.rint(inc)
        if reset_period > 0 and r > 0 and (row_hash(key, r, 2) % reset_period) == 0:
            v = inc
        else:
            v = np.float64(v) + inc
        vals[r] = v
    return ts, vals


def chunk_rows(rows, rows_per_chunk):
    out = []
    r = 0
    while r < rows:
        out.append(min(rows_per_chunk, rows - r))
        r += rows_per_chunk
    return out


def gen_hist_series(seed, gid, rows, nb, reset_period):
    """Cumulative bucket counts [rows, nb] of the histogram generator (hist_row in synth_kernels.cu): row r adds 1 + hash % 3 observations to
    bucket (r + gid) % nb; series with gid % reset_period == 0 restart at 5/8 of the rows."""
    key = series_key(seed, gid)
    cnt = np.zeros(nb, np.int64)
    out = np.zeros((rows, nb), np.int64)
    for r in range(rows):
        if reset_period > 0 and gid % reset_period == 0 and r == (rows * 5) // 8:
            cnt[:] = 0
        cnt[(r + gid) % nb] += 1 + int(row_hash(key, r, 7) % 3)
        out[r] = np.cumsum(cnt)
    return out
