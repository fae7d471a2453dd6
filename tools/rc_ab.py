"""A/B of two builds of the host range coder on streams of different entropy (host CPU only)."""
import ctypes as C, sys, time, numpy as np, hashlib
def bench(path):
    lib = C.CDLL(path)
    lib.linr_rc_encode_binary.restype = C.c_int64
    lib.linr_rc_encode_binary.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.linr_rc_decode_binary.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    n = 2_000_000
    for a in (0.03, 0.15, 1.0):
        rng = np.random.default_rng(1)
        p0 = np.clip(rng.beta(a, a, n), 1e-4, 1 - 1e-4)
        mid = (np.rint(p0 * 65534) + 1).astype(np.uint16)
        sym = (rng.random(n) > p0).astype(np.uint8)
        out = np.empty(n // 2 + 64, np.uint8); dec = np.empty(n, np.uint8)
        best_e = best_d = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); w = lib.linr_rc_encode_binary(mid.ctypes.data, sym.ctypes.data, n, out.ctypes.data, len(out)); t1 = time.perf_counter()
            lib.linr_rc_decode_binary(mid.ctypes.data, out.ctypes.data, w, dec.ctypes.data, n); t2 = time.perf_counter()
            best_e, best_d = min(best_e, t1 - t0), min(best_d, t2 - t1)
        assert (dec == sym).all()
        print(f"{path.split('/')[-1]:12s} beta({a}) {8*w/n:.3f} bit/sym  encode {n/best_e/1e6:6.0f} Msym/s  decode {n/best_d/1e6:6.0f} Msym/s  md5 {hashlib.md5(out[:w].tobytes()).hexdigest()[:8]}")
for p in sys.argv[1:]:
    bench(p)
