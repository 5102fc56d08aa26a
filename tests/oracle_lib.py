"""ctypes view of oracle/liborc.so — the CPU checker.  Test infrastructure only."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p = ctypes.POINTER(ctypes.c_uint32)
P = (1 << 31) - 1


def ptr(a):
    return a.ctypes.data_as(u32p)


def ptrs(arrs):
    return (u32p * max(1, len(arrs)))(*[ptr(a) for a in arrs])


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborc.so"))
        self.lib.orc_grind.restype = ctypes.c_uint64
        self.lib.orc_blake2s256.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]

    def twiddles(self, root_log):
        tw = np.empty(1 << root_log, dtype=np.uint32)
        itw = np.empty(1 << root_log, dtype=np.uint32)
        self.lib.orc_precompute_twiddles(root_log, ptr(tw), ptr(itw))
        return tw, itw

    def interpolate(self, v, root_log):
        c = np.ascontiguousarray(v, dtype=np.uint32).copy()
        self.lib.orc_interpolate(ptr(c), int(np.log2(c.size)), 1, root_log)
        return c

    def evaluate(self, coeffs, log_blowup, root_log):
        c = np.ascontiguousarray(coeffs, dtype=np.uint32)
        lg = int(np.log2(c.size))
        out = np.zeros(c.size << log_blowup, dtype=np.uint32)
        self.lib.orc_evaluate(ptr(c), lg, log_blowup, 1, root_log, ptr(out))
        return out

    def eval_at_point(self, coeffs, point8):
        c = np.ascontiguousarray(coeffs, dtype=np.uint32)
        p = np.ascontiguousarray(point8, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        self.lib.orc_eval_at_point(ptr(c), int(np.log2(c.size)), ptr(p), ptr(out))
        return out

    def domain_at(self, log, i):
        out = np.zeros(2, dtype=np.uint32)
        self.lib.orc_domain_at(log, i, ptr(out))
        return int(out[0]), int(out[1])

    def compress(self, h, m, t0=0, t1=0, f0=0, f1=0):
        hh = np.ascontiguousarray(h, dtype=np.uint32).copy()
        mm = np.ascontiguousarray(m, dtype=np.uint32)
        self.lib.orc_compress(ptr(hh), ptr(mm), ctypes.c_uint32(t0), ctypes.c_uint32(t1), ctypes.c_uint32(f0), ctypes.c_uint32(f1))
        return hh

    def blake2s256(self, data: bytes) -> bytes:
        out = (ctypes.c_uint8 * 32)()
        self.lib.orc_blake2s256(data, len(data), out)
        return bytes(out)

    def commit_on_layer(self, log, prev, cols):
        out = np.zeros(8 << log, dtype=np.uint32)
        self.lib.orc_commit_on_layer(log, ptr(prev) if prev is not None else None, ptrs(cols), len(cols), ptr(out))
        return out

    def merkle_commit(self, cols):
        """returns layers[k] for k = 0..max_log"""
        logs = np.array([int(np.log2(c.size)) for c in cols], dtype=np.uint32)
        max_log = int(logs.max())
        total = sum(8 << k for k in range(max_log + 1))
        buf = np.zeros(total, dtype=np.uint32)
        self.lib.orc_merkle_commit(ptrs(cols), ptr(logs), len(cols), ptr(buf))
        layers, off = {}, 0
        for k in range(max_log, -1, -1):
            layers[k] = buf[off:off + (8 << k)]
            off += 8 << k
        return [layers[k] for k in range(max_log + 1)]

    def bit_reverse(self, v):
        c = np.ascontiguousarray(v, dtype=np.uint32).copy()
        self.lib.orc_bit_reverse(ptr(c), int(np.log2(c.size)))
        return c

    def batch_inverse_m31(self, v):
        s = np.ascontiguousarray(v, dtype=np.uint32)
        d = np.zeros_like(s)
        self.lib.orc_batch_inverse_m31(ptr(s), ptr(d), ctypes.c_size_t(s.size))
        return d

    def batch_inverse_qm31(self, coords):
        d = [np.zeros_like(c) for c in coords]
        self.lib.orc_batch_inverse_qm31(ptrs(coords), ptrs(d), ctypes.c_size_t(coords[0].size))
        return d

    def fold_line(self, coords, alpha):
        lg = int(np.log2(coords[0].size))
        d = [np.zeros(coords[0].size // 2, dtype=np.uint32) for _ in range(4)]
        a = np.ascontiguousarray(alpha, dtype=np.uint32)
        self.lib.orc_fold_line(ptrs(coords), lg, ptr(a), ptrs(d))
        return d

    def fold_circle_into_line(self, dst, coords, alpha):
        lg = int(np.log2(coords[0].size))
        d = [np.ascontiguousarray(x, dtype=np.uint32).copy() for x in dst]
        a = np.ascontiguousarray(alpha, dtype=np.uint32)
        self.lib.orc_fold_circle_into_line(ptrs(coords), lg, ptr(a), ptrs(d))
        return d

    def prefix_sum_bitrev(self, v):
        c = np.ascontiguousarray(v, dtype=np.uint32).copy()
        self.lib.orc_prefix_sum_bitrev(ptr(c), int(np.log2(c.size)))
        return c

    def secure_powers(self, felt, n):
        f = np.ascontiguousarray(felt, dtype=np.uint32)
        out = np.zeros((n, 4), dtype=np.uint32)
        self.lib.orc_secure_powers(ptr(f), n, ptr(out))
        return out

    def accumulate_quotients(self, log, cols, alpha, bpts, bsizes, ecols, evals):
        out = [np.zeros(1 << log, dtype=np.uint32) for _ in range(4)]
        a = np.ascontiguousarray(alpha, dtype=np.uint32)
        bp, bs = np.ascontiguousarray(bpts, dtype=np.uint32), np.ascontiguousarray(bsizes, dtype=np.uint32)
        ec, ev = np.ascontiguousarray(ecols, dtype=np.uint32), np.ascontiguousarray(evals, dtype=np.uint32)
        self.lib.orc_accumulate_quotients(log, ptrs(cols), len(cols), ptr(a), ptr(bp), ptr(bs), ptr(ec), ptr(ev), bs.size, ptrs(out))
        return out

    def grind(self, digest, pow_bits):
        d = np.ascontiguousarray(digest, dtype=np.uint32)
        return int(self.lib.orc_grind(ptr(d), pow_bits))
